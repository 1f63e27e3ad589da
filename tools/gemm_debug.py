"""On-device diagnostics for the tcgen05 conv GEMM: structured operands whose wrong answers reveal WHICH
descriptor / layout assumption is off.  Prints, never asserts.  Usage: python tools/gemm_debug.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from voice100_b200 import kernels as K

DEV = "cuda"


def ncw(x):
    B, C, T = x.shape
    out = K.empty_ncw(B, C, T, x.device)
    out.data.zero_()
    out.data[:, :, :T] = x.to(torch.bfloat16)
    return out


def report(tag, got, ref):
    err = (got.double() - ref.double()).abs()
    print(f"[{tag}] max_abs_err={float(err.max()):.4g} ref_max={float(ref.abs().max()):.4g} "
          f"frac_bad={(err > 0.05 * ref.abs().max()).double().mean():.4f}")
    if float(err.max()) > 0.05 * float(ref.abs().max()):
        bad = (err > 0.05 * ref.abs().max()).nonzero()
        print("   first bad idx:", bad[:6].tolist())
        b, c, t = bad[0].tolist()
        print("   got row :", [round(float(v), 3) for v in got[b, c, t:t + 10]])
        print("   ref row :", [round(float(v), 3) for v in ref[b, c, t:t + 10]])
        print("   got col :", [round(float(v), 3) for v in got[b, c:c + 10, t]])
        print("   ref col :", [round(float(v), 3) for v in ref[b, c:c + 10, t]])
        # per 8-row / 64-col block error map of the first tile
        e = err[0, :128, :256]
        if e.shape[0] >= 8 and e.shape[1] >= 64:
            blocks = e[: e.shape[0] // 8 * 8, : e.shape[1] // 64 * 64].reshape(e.shape[0] // 8, 8, e.shape[1] // 64, 64)
            print("   block max err (rows/8 x cols/64):")
            print(blocks.amax(dim=(1, 3)).cpu().numpy().round(2))


def main():
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    zero = lambda n: torch.zeros(n, device=DEV)
    # 1. identity weights: y must equal x (tests B-operand layout + TMEM/epilogue/TMA-store mapping)
    for C, T in ((64, 64), (128, 256), (128, 300)):
        x = torch.randn(1, C, T, device=DEV)
        W = torch.eye(C, device=DEV).to(torch.bfloat16)
        xn = ncw(x)
        y = K.conv1x1(xn, W, None, zero(C), 0)
        torch.cuda.synchronize()
        report(f"identity C={C} T={T}", y.valid().float(), xn.valid().float())
    # 2. x = time index pattern, W = one-hot rows: y[co][t] = x[perm[co]][t]
    C, T = 128, 256
    x = (torch.arange(C, device=DEV)[:, None] * 1.0 + torch.arange(T, device=DEV)[None, :] / 256.0)[None]
    perm = torch.randperm(C, device=DEV)
    W = torch.zeros(C, C, device=DEV)
    W[torch.arange(C), perm] = 1
    xn = ncw(x)
    y = K.conv1x1(xn, W.to(torch.bfloat16), None, zero(C), 0)
    torch.cuda.synchronize()
    report("permutation", y.valid().float(), xn.valid().float()[:, perm])
    # 3. random, several shapes
    for (B, Ci, Co, T) in ((1, 64, 128, 64), (1, 256, 128, 256), (2, 64, 256, 1501), (3, 1024, 256, 751), (40, 128, 384, 520)):
        x = torch.randn(B, Ci, T, device=DEV)
        W = (torch.randn(Co, Ci, device=DEV) / Ci ** 0.5).to(torch.bfloat16)
        xn = ncw(x)
        y = K.conv1x1(xn, W, None, zero(Co), 0)
        torch.cuda.synchronize()
        ref = torch.einsum("oc,bct->bot", W.float(), xn.valid().float())
        report(f"random B={B} {Ci}->{Co} T={T}", y.valid().float(), ref)
    # 4. fp32-out head and residual path
    x = torch.randn(2, 256, 300, device=DEV)
    W = (torch.randn(29, 256, device=DEV) / 16).to(torch.bfloat16)
    xn = ncw(x)
    y = K.conv1x1_f32(xn, W, zero(29))
    torch.cuda.synchronize()
    report("f32out 256->29", y.valid(), torch.einsum("oc,bct->bot", W.float(), xn.valid().float()))
    x = torch.randn(2, 128, 300, device=DEV)
    r = torch.randn(2, 128, 300, device=DEV)
    W = (torch.randn(128, 128, device=DEV) / 11).to(torch.bfloat16)
    xn, rn = ncw(x), ncw(r)
    y = K.conv1x1(xn, W, None, zero(128), 0, rn)
    torch.cuda.synchronize()
    report("residual", y.valid().float(), torch.einsum("oc,bct->bot", W.float(), xn.valid().float()) + rn.valid().float())
    print("gemm_debug done")


if __name__ == "__main__":
    main()
