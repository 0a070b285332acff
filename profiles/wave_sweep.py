#!/usr/bin/env python
"""Step time against batch size around whole waves of CTAs (8 CTAs/SM x 148 SMs x 16 envs = 18944 envs
per wave): does the 65536-env headline workload pay for a partially filled last wave?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tune_step import time_cfg  # noqa: E402

if __name__ == "__main__":
    for B in (18944, 37888, 47360, 56832, 60000, 65536, 71040, 75776, 94720, 131072, 262144):
        ms = time_cfg("harvest", B, steps=200, warm=100)
        print("B=%7d  %.4f ms/step  %.3f G agent-steps/s  (%.2f waves)" % (B, ms, B * 5 / ms / 1e6, B / 18944.0), flush=True)
