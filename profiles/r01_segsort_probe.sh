#!/bin/bash
# quick probe of the experimental per-read arc sort: parity with the reference on two small sets, timing on 100 K reads
cd /root/repo
python -c "
from miniasm_b200 import synth
synth.generate('chaos','/tmp/c.paf'); synth.generate('-n 1500 -l 3000 -L 12000 -c 400 -j 30 -s 77','/tmp/d.paf'); synth.generate('c2_100k','/tmp/e.paf')"
for f in c d; do
  oracle/_ref/miniasm_ref /tmp/$f.paf > /tmp/$f.ref 2>/dev/null
  MAB_SG_SEGSORT=1 timeout 60 miniasm_b200/miniasm-b200 /tmp/$f.paf > /tmp/$f.out 2>/tmp/$f.err; echo "rc=$?"
  cmp -s /tmp/$f.ref /tmp/$f.out && echo SAME_$f || { echo DIFF_$f; tail -3 /tmp/$f.err; }
done
MAB_TRACE=1 timeout 60 miniasm_b200/miniasm-b200 /tmp/e.paf > /tmp/e0.out 2>/tmp/e0.err
MAB_TRACE=1 MAB_SG_SEGSORT=1 timeout 60 miniasm_b200/miniasm-b200 /tmp/e.paf > /tmp/e1.out 2>/tmp/e1.err; echo "rc=$?"
cmp -s /tmp/e0.out /tmp/e1.out && echo SAME_e || echo DIFF_e
grep -i "sg_gen\|layout" /tmp/e0.err | head -5; echo ---; grep -i "sg_gen\|layout" /tmp/e1.err | head -5
