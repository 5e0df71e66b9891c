#!/bin/bash
# On the GPU box: frame_timeline / call_timeline for every variants/*.so in one run (A/B on one box), alternating twice.
LIB=orb-slam-birdview_b200/liborbb200.so
cp $LIB /tmp/orig.so
for rep in 1 2; do
  for v in variants/*.so; do cp $v $LIB; echo "$(basename $v) $(tools/ubench/frame_timeline 200)"; done
done
cp /tmp/orig.so $LIB
