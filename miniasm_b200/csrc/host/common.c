/* common.c -- option defaults and the verbosity global (reference: common.c:3-23). */
#include "miniasm_b200.h"

int ma_verbose = 3;

void ma_opt_init(ma_opt_t *o)
{
	o->min_span = 2000, o->min_match = 100, o->min_dp = 3, o->min_iden = .05f;
	o->max_hang = 1000, o->min_ovlp = o->min_span, o->int_frac = .8f;
	o->gap_fuzz = 1000, o->n_rounds = 2, o->bub_dist = 50000, o->max_ext = 4;
	o->min_ovlp_drop_ratio = .5f, o->max_ovlp_drop_ratio = .7f, o->final_ovlp_drop_ratio = .8f;
}
