import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moonshine_b200 import api
lib = api.load_library()
MB = 1 << 20
print("src     stage  stages nsub  grid   GB/s/SM   TB/s total")
for shared in (0, 1):
    for grid in (148, 1):
        for stage_kb, stages, nsub in [(32, 6, 1), (32, 6, -2), (32, 6, -3), (32, 6, -6), (16, 12, 1), (16, 12, -4),
                                       (16, 12, -8), (8, 24, -8), (64, 3, 1), (64, 3, -3), (32, 3, -3), (32, 2, -2)]:
            per = 8 * MB
            ms = lib.moonshine_b200_test_ring_bandwidth(per, stage_kb * 1024, stages, nsub, shared, grid)
            if ms <= 0:
                continue
            gbs = per / (ms * 1e-3) / 1e9
            print(f"{'L2' if shared else 'HBM':6s} {stage_kb:4d}K {stages:6d} {nsub:4d} {grid:5d} {gbs:9.1f} {gbs * grid / 1000:9.2f}")
