import sys, json
sys.path.insert(0,'/root/repo')
import torch, bench
import orb_slam_birdview_b200 as pkg
print(json.dumps(bench.leg_knn2(torch, pkg, 0, reps=20)))
