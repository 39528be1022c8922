import sys, os, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moonshine_b200 import api
lib = api.load_library()
shapes = [(106752, 1248, 416, 0), (106752, 1664, 416, 1), (106752, 416, 1664, 0), (13344, 576, 2016, 1), (13344, 288, 288, 0), (415, 415, 52, 0)]
only = sys.argv[1] if len(sys.argv) > 1 else None
for (M, N, K, act) in shapes:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); C = torch.zeros(M, N, device="cuda"); b = torch.randn(N, device="cuda")
    for impl, name in ((1, "simt"), (2, "tc")):
        if only and name != only: continue
        for _ in range(2):
            lib.moonshine_b200_test_gemm(A.data_ptr(), W.data_ptr(), C.data_ptr(), M, N, K, K, K, N, b.data_ptr(), act, 0, impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            lib.moonshine_b200_test_gemm(A.data_ptr(), W.data_ptr(), C.data_ptr(), M, N, K, K, K, N, b.data_ptr(), act, 0, impl)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name:5s} M={M} N={N} K={K} act={act}: {ms:8.3f} ms  {2*M*N*K/ms/1e9:8.1f} TFLOP/s", flush=True)
