#!/bin/bash
# On the GPU box: the drop-in path with each latency measure of round 2c switched off in turn, all in ONE run on one box
# (box-to-box variation is larger than some of the effects).  Output: one JSON line per (probe, knob set).
#   tools/latency_ab.sh [reps]
R=${1:-200}
for k in "" "ORBB200_NO_HOST_GRAPH=1" "ORBB200_NO_PDL=1" "ORBB200_NO_SPLIT=1" "ORBB200_NO_OCTREE_SMEM=1" \
         "ORBB200_NO_HOST_GRAPH=1 ORBB200_NO_PDL=1 ORBB200_NO_SPLIT=1 ORBB200_NO_OCTREE_SMEM=1"; do
  for shape in "752 480 1000" "1241 376 2000"; do
    echo "{\"probe\": \"call_timeline\", \"knobs\": \"$k\", \"result\": $(env $k tools/ubench/call_timeline $shape $R)}"
  done
done
for k in "" "ORBB200_NO_FRAME_GRAPH=1" "ORBB200_SELECT_TIERS=1" "ORBB200_SUBPIX_GENERIC=1" "ORBB200_NO_WARP_CANDS=1" \
         "ORBB200_NO_FRAME_GRAPH=1 ORBB200_SELECT_TIERS=1 ORBB200_SUBPIX_GENERIC=1 ORBB200_NO_WARP_CANDS=1 ORBB200_NO_HOST_GRAPH=1 ORBB200_NO_PDL=1 ORBB200_NO_SPLIT=1 ORBB200_NO_OCTREE_SMEM=1"; do
  echo "{\"probe\": \"frame_timeline\", \"knobs\": \"$k\", \"result\": $(env $k tools/ubench/frame_timeline $R)}"
done
for k in "" "ORBB200_NO_WARP_CANDS=1" "ORBB200_NO_STAGED_MATCH=1" "ORBB200_NO_WARP_CANDS=1 ORBB200_NO_STAGED_MATCH=1"; do
  echo "{\"probe\": \"matcher_latency\", \"knobs\": \"$k\", \"result\": $(env $k python tools/matcher_latency.py $R | tail -1)}"
done
