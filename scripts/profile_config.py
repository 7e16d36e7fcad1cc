"""Run one BASELINE config at a fixed ensemble size (for ncu). usage: profile_config.py <3|4a|4b|5c> <B>"""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
import bench_configs as bc
name, B = sys.argv[1], int(sys.argv[2])
make = {"3": bc.config3, "4a": bc.config4a, "4b": bc.config4b, "5c": lambda B: bc.config5(B, d=256, constraint="ts1")}[name]
run = make(B)
run(); torch.cuda.synchronize()
run(); torch.cuda.synchronize()
