"""attention_tc2_kernel (softmax + epilogue warpgroups) against the first-generation tcgen05 kernel
and a torch fp32 reference, on the bench shape and on ragged / tiny cases.  Each kernel generation runs in its own
process (FC_ATTENTION is read once per process)."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [(3, 197, 12), (256, 197, 12), (1, 193, 1), (5, 208, 2), (7, 200, 16)]


def run(tag):
    from fitclip_b200 import ops
    dev = torch.device("cuda:0")
    outs = {}
    for seqs, L, heads in CASES:
        g = torch.Generator(device=dev).manual_seed(seqs * 1000 + L)
        qkv = (torch.randn(seqs * L, 3 * heads * 64, device=dev, generator=g) * 1.5).bfloat16()
        qkv[:, :heads * 64] *= 2.0  # sharper rows
        out = ops.attention_bf16(qkv, seqs, L, heads, False)
        torch.cuda.synchronize()
        q, k, v = (qkv.float().view(seqs, L, 3, heads, 64)[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        ref = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v
        ref = ref.permute(0, 2, 1, 3).reshape(seqs * L, heads * 64)
        err = (out.float() - ref).abs().max().item()
        print(f"[{tag}] seqs={seqs} L={L} heads={heads}: max abs err vs fp32 torch {err:.3e} (ref max {ref.abs().max().item():.2f})", flush=True)
        assert err < 3e-2, err
        outs[(seqs, L, heads)] = out.cpu()
    torch.save(outs, f"/tmp/att_{tag}.pt")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
        sys.exit(0)
    for tag, env in (("tc2", {}), ("tc1", {"FC_ATTENTION": "tc1"})):
        r = subprocess.run([sys.executable, __file__, tag], env={**os.environ, **env}, timeout=120)
        print(f"{tag}: exit {r.returncode}", flush=True)
        if r.returncode:
            sys.exit(r.returncode)
    b = torch.load("/tmp/att_tc1.pt")
    for tag in ("tc2",):
        a = torch.load(f"/tmp/att_{tag}.pt")
        for key in a:
            d = (a[key].float() - b[key].float()).abs().max().item()
            print(f"{tag} vs tc1 {key}: max abs diff {d:.3e}")
            assert d < 7e-2  # two bf16 ulps at the largest outputs (|o| < 8: ulp 0.03125)
    print("attention_check OK")
