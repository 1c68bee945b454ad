"""Pipeline timeline of one attention-backward CTA (debug build only: `make -C sam3_lora_b200/csrc TRACE=1`).

    SAM3B_LIB=sam3_lora_b200/libsam3b_trace.so python tools/attn_trace.py 576 72 16

Prints, for the traced dK/dV CTA, per 64-query block the clock deltas between the pipeline events stamped in
csrc/attn_bwd.cu (TRACE slots): compute warp 0 and the MMA-issuing thread."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("SAM3B_LIB", str(ROOT / "sam3_lora_b200" / "libsam3b_trace.so"))

import torch  # noqa: E402

from sam3_lora_b200 import _lib as L  # noqa: E402


def main():
    Ls, segs, heads = (int(a) for a in sys.argv[1:4])
    dev = "cuda"
    dt = torch.float16
    D = heads * 64
    T = Ls * segs
    qkv = (torch.randn(T, 3 * D + 64, device=dev) * 0.5).to(dt)
    O = torch.zeros(T, D + 64, device=dev, dtype=dt)
    lse2 = torch.zeros(heads, T, device=dev)
    dO = (torch.randn(T, D, device=dev) * 0.5).to(dt)
    delta = torch.zeros(heads, T, device=dev)
    dqkv = torch.zeros(T, 3 * D + 64, device=dev, dtype=dt)
    tab = torch.zeros(Ls, 32, 2, device=dev)
    tab[..., 0] = 1.0
    L.attention_fwd(qkv, Ls, D, heads, O, lse2)
    lib = L.load()
    lib.sam3b_debug_trace_read.argtypes = [C.c_void_p, C.c_int]
    for _ in range(2):
        L.attention_bwd(qkv, Ls, D, heads, O, lse2, dO, delta, dqkv, tab, Ls)
    torch.cuda.synchronize()
    lib.sam3b_debug_trace_clear()
    L.attention_bwd(qkv, Ls, D, heads, O, lse2, dO, delta, dqkv, tab, Ls)
    torch.cuda.synchronize()
    n = 16384
    buf = (C.c_ulonglong * n)()
    lib.sam3b_debug_trace_read(buf, n)
    t = list(buf)
    nb = (Ls + 63) // 64
    t0 = t[0]
    hdr = {"setup_done": t[1] - t0, "kv_ready(mma)": t[2] - t0, "last_acc_done": t[3] - t0, "epilogue_end": t[4] - t0}
    print(json.dumps({"case": f"dkdv_{Ls}_{segs}_{heads}", "blocks": nb, "header_clk": hdr}))
    rows = []
    for j in range(nb):
        b = 64 + j * 16
        e = t[b:b + 16]
        rows.append({
            "j": j,
            "c_start": e[0] - t0,
            "c_wait_sdp": e[1] - e[0], "c_ld": e[2] - e[1], "c_math": e[3] - e[2], "c_wait_acc": e[4] - e[3], "c_store": e[5] - e[4],
            "m_start": e[8] - t0, "m_issue_sdp(wait+issue)": e[9] - e[8], "m_wait_pds": e[10] - e[9], "m_issue_acc": e[11] - e[10],
        })
    show = rows if nb <= 12 else rows[:4] + rows[nb // 2:nb // 2 + 3] + rows[-3:]
    for r in show:
        print(json.dumps(r))
    if nb > 2:
        per = (rows[-1]["c_start"] - rows[1]["c_start"]) / (nb - 2)
        print(json.dumps({"steady_clk_per_block": per}))


if __name__ == "__main__":
    main()
