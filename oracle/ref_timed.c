/* ref_timed.c -- ORACLE / TEST INFRASTRUCTURE, not part of the product.
 *
 * A timing harness around the UNMODIFIED reference objects (hit.o asg.o asm.o paf.o sdict.o
 * sys.o common.o, compiled by oracle/Makefile from the sources where they lie).  It repeats the
 * default step order of the reference driver (main.c:108-199, default options common.c:5-23,
 * no -R/-1/-2/-S) and wraps every library call with a wall clock, so bench.py can quote the
 * reference's own per-function CPU time (e.g. asg_arc_del_trans) next to the GPU number.
 *
 * usage: miniasm_ref_timed in.paf > out.gfa      (timings: one JSON object on stderr, tag "[T]")
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "miniasm.h"
#include "sys.h"

static double now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

#define NT 32
static const char *t_name[NT];
static double t_sec[NT];
static int n_t;

static void rec(const char *name, double t0)
{
	double dt = now() - t0;
	int i;
	for (i = 0; i < n_t; ++i)
		if (strcmp(t_name[i], name) == 0) { t_sec[i] += dt; return; }
	if (n_t < NT) t_name[n_t] = name, t_sec[n_t++] = dt;
}

#define TIMED(name, stmt) do { double t0_ = now(); stmt; rec(name, t0_); } while (0)

int main(int argc, char *argv[])
{
	ma_opt_t opt;
	sdict_t *d;
	ma_sub_t *sub, *sub2;
	ma_hit_t *hit;
	size_t n_hits, n_lines_hits;
	float cov;
	asg_t *sg;
	ma_ug_t *ug;
	uint32_t n_arc_in;
	int i, changed;
	double t_all = now();

	if (argc < 2) { fprintf(stderr, "usage: %s in.paf > out.gfa\n", argv[0]); return 1; }
	ma_opt_init(&opt);
	opt.min_ovlp = opt.min_span;
	sys_init();
	d = sd_init();

	TIMED("ma_hit_read", hit = ma_hit_read(argv[1], opt.min_span, opt.min_match, d, &n_hits, 1, 0));
	n_lines_hits = n_hits;
	TIMED("ma_hit_sub", sub = ma_hit_sub(opt.min_dp, opt.min_iden, 0, n_hits, hit, d->n_seq));
	TIMED("ma_hit_cut", n_hits = ma_hit_cut(sub, opt.min_span, n_hits, hit));
	TIMED("ma_hit_flt", n_hits = ma_hit_flt(sub, opt.max_hang * 1.5, opt.min_ovlp * .5, n_hits, hit, &cov));
	TIMED("ma_hit_sub", sub2 = ma_hit_sub(opt.min_dp, opt.min_iden, opt.min_span / 2, n_hits, hit, d->n_seq));
	TIMED("ma_hit_cut", n_hits = ma_hit_cut(sub2, opt.min_span, n_hits, hit));
	TIMED("ma_sub_merge", ma_sub_merge(d->n_seq, sub, sub2));
	free(sub2);
	TIMED("ma_hit_contained", n_hits = ma_hit_contained(&opt, d, sub, n_hits, hit));
	hit = (ma_hit_t*)realloc(hit, n_hits * sizeof(ma_hit_t));

	TIMED("ma_sg_gen", sg = ma_sg_gen(&opt, d, sub, n_hits, hit));
	n_arc_in = sg->n_arc;
	TIMED("asg_arc_del_trans", asg_arc_del_trans(sg, opt.gap_fuzz));
	TIMED("asg_cut_tip", asg_cut_tip(sg, opt.max_ext));
	TIMED("asg_pop_bubble", asg_pop_bubble(sg, opt.bub_dist));
	for (i = 0; i <= opt.n_rounds; ++i) {
		float r = opt.min_ovlp_drop_ratio + (opt.max_ovlp_drop_ratio - opt.min_ovlp_drop_ratio) / opt.n_rounds * i;
		TIMED("asg_arc_del_short", changed = asg_arc_del_short(sg, r));
		if (changed) {
			TIMED("asg_cut_tip", asg_cut_tip(sg, opt.max_ext));
			TIMED("asg_pop_bubble", asg_pop_bubble(sg, opt.bub_dist));
		}
	}
	TIMED("asg_cut_internal", asg_cut_internal(sg, 1));
	TIMED("asg_cut_biloop", asg_cut_biloop(sg, opt.max_ext));
	TIMED("asg_cut_tip", asg_cut_tip(sg, opt.max_ext));
	TIMED("asg_pop_bubble", asg_pop_bubble(sg, opt.bub_dist));
	TIMED("asg_arc_del_short", changed = asg_arc_del_short(sg, opt.final_ovlp_drop_ratio));
	if (changed) {
		TIMED("asg_cut_tip", asg_cut_tip(sg, opt.max_ext));
		TIMED("asg_pop_bubble", asg_pop_bubble(sg, opt.bub_dist));
	}
	TIMED("ma_ug_gen", ug = ma_ug_gen(sg));
	if (argc > 2) TIMED("ma_ug_seq", ma_ug_seq(ug, d, sub, argv[2]));
	TIMED("ma_ug_print", ma_ug_print(ug, d, sub, stdout));
	fflush(stdout);

	fprintf(stderr, "[T] {\"total\": %.6f, \"n_hits_stored\": %lu, \"n_arc_del_trans_in\": %u", now() - t_all,
			(unsigned long)n_lines_hits, n_arc_in);
	for (i = 0; i < n_t; ++i) fprintf(stderr, ", \"%s\": %.6f", t_name[i], t_sec[i]);
	fprintf(stderr, "}\n");

	asg_destroy(sg); ma_ug_destroy(ug);
	free(sub); free(hit); sd_destroy(d);
	return 0;
}
