#!/bin/bash
# On the GPU box: call_timeline for every variants/*.so in one run (A/B on one box), alternating twice.
LIB=orb-slam-birdview_b200/liborbb200.so
cp $LIB /tmp/orig.so
for rep in 1 2; do
  for v in variants/*.so; do cp $v $LIB; echo "$(basename $v) $(tools/ubench/call_timeline 752 480 1000 300 | cut -c1-150) $(tools/ubench/call_timeline 1241 376 2000 300 | cut -c1-150)"; done
done
cp /tmp/orig.so $LIB
