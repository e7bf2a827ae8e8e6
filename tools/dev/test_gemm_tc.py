import sys, torch
sys.path.insert(0, '.')
from codlad_b200 import _native as N
lib = N.lib()
torch.manual_seed(0)
def run(M, Nn, K, a_kc, b_kc, acc, lda_pad=0):
    A = torch.randn((M, K) if a_kc else (K, M), device="cuda")
    B = torch.randn((Nn, K) if b_kc else (K, Nn), device="cuda")
    C0 = torch.randn(M, Nn, device="cuda")
    ref = (A.double() if a_kc else A.double().t()) @ (B.double().t() if b_kc else B.double()) + (C0.double() if acc else 0)
    out = {}
    for mode in (0, 1):
        N.check(lib.cb2t_set_gemm_mode(mode))
        C = C0.clone()
        N.check(lib.cb2t_gemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, Nn, K, A.shape[1], B.shape[1], Nn, a_kc, b_kc, acc, N.stream_ptr()))
        torch.cuda.synchronize()
        out[mode] = float((C.double() - ref).abs().max() / ref.abs().max())
    # timing of both modes
    t = {}
    for mode in (0, 1):
        N.check(lib.cb2t_set_gemm_mode(mode))
        C = C0.clone()
        for _ in range(2): lib.cb2t_gemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, Nn, K, A.shape[1], B.shape[1], Nn, a_kc, b_kc, 0, N.stream_ptr())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): lib.cb2t_gemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, Nn, K, A.shape[1], B.shape[1], Nn, a_kc, b_kc, 0, N.stream_ptr())
        b.record(); torch.cuda.synchronize()
        t[mode] = a.elapsed_time(b) / 5
    fl = 2.0 * M * Nn * K
    print(f"M={M} N={Nn} K={K} a_kc={a_kc} b_kc={b_kc} acc={acc}: err simt {out[0]:.2e} tf32 {out[1]:.2e} | ms simt {t[0]:.3f} ({fl/t[0]/1e9:.1f} TF/s) tf32 {t[1]:.3f} ({fl/t[1]/1e9:.1f} TF/s)")
N.check(lib.cb2t_set_gemm_mode(0))
import os
cases = [(4096, 128, 128, 1, 1, 0), (4096, 128, 128, 1, 1, 1), (5000, 512, 128, 1, 1, 0), (3001, 128, 512, 1, 1, 0), (200000, 128, 128, 1, 1, 0),
             (2000, 1152, 128, 1, 1, 0), (4096, 128, 152, 1, 1, 0),
             (128, 128, 16384, 0, 0, 0), (128, 128, 16384, 0, 0, 1), (512, 128, 100000, 0, 0, 0), (128, 512, 100000, 0, 0, 1), (128, 128, 1900000, 0, 0, 0)]
for args in cases:
    run(*args)
N.check(lib.cb2t_set_gemm_mode(0))
