"""Step-by-step driver of the miniasm C ABI, mirroring main.c:108-199 of the reference.

``Pipeline(lib, paf)`` works with any library that exports the reference seam (the CUDA product, the
unmodified reference, the oracle port), which is what lets the parity tests run the same steps on
both sides and compare the host-visible state after each of them.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import ARC_DT, HIT_DT, SUB_DT


class Pipeline:
    def __init__(self, lib, paf_path, opt=None, bi_dir=1):
        self.lib = lib
        self.opt = opt if opt is not None else lib.default_opt()
        self.paf = paf_path.encode() if isinstance(paf_path, str) else paf_path
        self.bi_dir = bi_dir
        self.d = None
        self.hits = None
        self.n_hits = 0
        self.sub = None
        self.sg = None
        self.ug = None
        self.cov = C.c_float(40.0)

    # ---- Step 1 ---------------------------------------------------------------------------------
    def read(self):
        lib, o = self.lib, self.opt
        self.d = lib.sd_init()
        n = C.c_size_t(0)
        self.hits = lib.ma_hit_read(self.paf, o.min_span, o.min_match, self.d, C.byref(n), self.bi_dir, None)
        self.n_hits = n.value
        return self

    def adopt(self, other):
        """Start from a copy of another pipeline's host state (hits, sub, dictionary lengths and names)."""
        lib = self.lib
        self.d = lib.sd_init()
        od = other.d.contents
        for i in range(od.n_seq):
            lib.sd_put(self.d, od.seq[i].name, od.seq[i].len)
            self.d.contents.seq[i].aux_del = od.seq[i].aux_del
        self.hits = capi.c_malloc_copy(other.hits_np())
        self.n_hits = other.n_hits
        self.sub = capi.c_malloc_copy(other.sub_np()) if other.sub else None
        return self

    # ---- Steps 2-3 (hit.c:109-256) --------------------------------------------------------------
    def sub1(self):
        lib, o = self.lib, self.opt
        self.sub = lib.ma_hit_sub(o.min_dp, o.min_iden, 0, self.n_hits, self.hits, self.d.contents.n_seq)
        return self

    def cut(self, sub=None):
        self.n_hits = self.lib.ma_hit_cut(sub if sub is not None else self.sub, self.opt.min_span, self.n_hits, self.hits)
        return self

    def flt(self):
        o = self.opt
        self.n_hits = self.lib.ma_hit_flt(self.sub, int(o.max_hang * 1.5), int(o.min_ovlp * .5), self.n_hits, self.hits,
                                          C.byref(self.cov))
        return self

    def sub2_cut_merge(self):
        lib, o = self.lib, self.opt
        sub2 = lib.ma_hit_sub(o.min_dp, o.min_iden, o.min_span // 2, self.n_hits, self.hits, self.d.contents.n_seq)
        self.n_hits = lib.ma_hit_cut(sub2, o.min_span, self.n_hits, self.hits)
        lib.ma_sub_merge(self.d.contents.n_seq, self.sub, sub2)
        capi.c_free(sub2)
        return self

    def contained(self):
        self.n_hits = self.lib.ma_hit_contained(C.byref(self.opt), self.d, self.sub, self.n_hits, self.hits)
        return self

    def select(self):
        return self.sub1().cut().flt().sub2_cut_merge().contained()

    # ---- Step 4 (asm.c:9-39, asg.c) -------------------------------------------------------------
    def sg_gen(self):
        self.sg = self.lib.ma_sg_gen(C.byref(self.opt), self.d, self.sub, self.n_hits, self.hits)
        return self

    def clean(self, upto=11):
        """main.c:156-188; `upto` has the meaning of the reference's -S stage."""
        lib, o, g = self.lib, self.opt, self.sg
        if upto >= 6:
            lib.asg_arc_del_trans(g, o.gap_fuzz)
        if upto >= 7:
            lib.asg_cut_tip(g, o.max_ext)
            lib.asg_pop_bubble(g, o.bub_dist)
        if upto >= 9:
            for i in range(o.n_rounds + 1):
                r = np.float32(o.min_ovlp_drop_ratio) + (np.float32(o.max_ovlp_drop_ratio) - np.float32(o.min_ovlp_drop_ratio)) \
                    / np.float32(o.n_rounds) * np.float32(i)
                if lib.asg_arc_del_short(g, float(r)) != 0:
                    lib.asg_cut_tip(g, o.max_ext)
                    lib.asg_pop_bubble(g, o.bub_dist)
        if upto >= 10:
            lib.asg_cut_internal(g, 1)
            lib.asg_cut_biloop(g, o.max_ext)
            lib.asg_cut_tip(g, o.max_ext)
            lib.asg_pop_bubble(g, o.bub_dist)
        if upto >= 11:
            if lib.asg_arc_del_short(g, o.final_ovlp_drop_ratio) != 0:
                lib.asg_cut_tip(g, o.max_ext)
                lib.asg_pop_bubble(g, o.bub_dist)
        return self

    # ---- Step 5 ---------------------------------------------------------------------------------
    def ug_gen(self):
        self.ug = self.lib.ma_ug_gen(self.sg)
        return self

    def gfa(self, reads=None):
        if reads:
            self.lib.ma_ug_seq(self.ug, self.d, self.sub, reads.encode())
        return self.lib.print_to_string("ma_ug_print", self.ug, self.d, self.sub)

    def sg_text(self):
        return self.lib.print_to_string("ma_sg_print", self.sg, self.d, self.sub)

    def run_all(self, reads=None):
        return self.read().select().sg_gen().clean().ug_gen().gfa(reads)

    # ---- host state as numpy --------------------------------------------------------------------
    def hits_np(self):
        return capi.np_from_ptr(self.hits, self.n_hits, HIT_DT)

    def sub_np(self):
        return capi.np_from_ptr(self.sub, self.d.contents.n_seq, SUB_DT)

    def names(self):
        d = self.d.contents
        return [d.seq[i].name for i in range(d.n_seq)]

    def seq_lens(self):
        d = self.d.contents
        return np.array([d.seq[i].len for i in range(d.n_seq)], dtype=np.uint32)

    def graph_np(self):
        return self.lib.read_graph(self.sg)

    def free(self):
        lib = self.lib
        if self.ug:
            lib.ma_ug_destroy(self.ug)
        if self.sg:
            lib.asg_destroy(self.sg)
        if self.sub:
            capi.c_free(self.sub)
        if self.hits:
            capi.c_free(self.hits)
        if self.d:
            lib.sd_destroy(self.d)
        self.ug = self.sg = self.sub = self.hits = self.d = None


def canon_arcs(arcs):
    """Arcs in a canonical total order (ties of the sort key broken by the remaining fields)."""
    a = np.asarray(arcs, dtype=ARC_DT)
    order = np.lexsort((a["ol_del"], a["v"], a["ul"]))
    return a[order]


def gfa_canon(text):
    """GFA text as the comparison form BASELINE.json asks for: S/L/a/x lines as a sorted multiset."""
    if isinstance(text, bytes):
        text = text.decode()
    return sorted(text.splitlines())
