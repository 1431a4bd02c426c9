"""Per-frame pipeline timing (config 3): python profiles/time_pipeline.py [n_frames]"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vdo_slam_b200 import capi
ctx = capi.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
print(json.dumps(bench.frames_per_second(ctx, n_frames=n), indent=1))
