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
    torch.cuda.synchronize()
    # ---- forward timeline (CTA x=200, head 3): softmax thread 0 and the MMA thread ----
    nbf = (Ls + 63) // 64
    fb = (C.c_ulonglong * 4096)()
    lib.sam3b_debug_trace_read_fwd.argtypes = [C.c_void_p, C.c_int]
    lib.sam3b_debug_trace_read_fwd(fb, 4096)
    ft = list(fb)
    f0 = ft[64]
    print(json.dumps({"case": f"fwd_{Ls}_{segs}_{heads}", "blocks": nbf}))
    frows = []
    for j in range(min(nbf, 96)):
        e = ft[64 + j * 8:64 + j * 8 + 8]
        m = ft[1024 + j * 4:1024 + j * 4 + 4]
        frows.append({"j": j, "c_start": e[0] - f0, "c_wait_s": e[1] - e[0], "c_ld": e[2] - e[1], "c_max": e[3] - e[2], "c_exp_pack": e[4] - e[3],
                      "c_st_arrive": e[5] - e[4], "m_s_issued": m[0] - f0, "m_pv_issued": m[1] - f0})
    fshow = frows if len(frows) <= 12 else frows[:3] + frows[len(frows) // 2:len(frows) // 2 + 3] + frows[-2:]
    for r in fshow:
        print(json.dumps(r))
    if len(frows) > 4:
        print(json.dumps({"fwd_steady_clk_per_block": (frows[-1]["c_start"] - frows[1]["c_start"]) / (len(frows) - 2)}))
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
    # slots (csrc/attn_bwd.cu), items 2 and 3 of CTA 70 (item 3 at +4096): compute thread 0 / 256 (group 0 / 1): 64 + j*8 +
    # {0 start, 1 S^T/dP^T ready, 2 in registers, 3 math+pack done, 4 P^T/dS^T buffer free, 5 stored + signalled};
    # MMA threads: 1024 + j*4 + {0 S^T/dP^T(j) issued, 1 dV/dK(j) issued};
    # item transition per group g: 8 + g*8 + {0 block loop left, 1 kv_free seen, 2 next K/V parked, 3 all_done seen, 4 epilogue stored}
    def rows_of(base):
        out = []
        for j in range(min(nb, 64)):
            e = t[base + 64 + j * 8:base + 64 + j * 8 + 8]
            m = t[base + 1024 + j * 4:base + 1024 + j * 4 + 4]
            out.append((e, m))
        return out
    it2, it3 = rows_of(0), rows_of(4096)
    t0 = min(x for e, m in it2 for x in list(e[:6]) + list(m[:2]) if x)
    print(json.dumps({"case": f"dkdv_{Ls}_{segs}_{heads}", "blocks": nb, "note": "items 2 and 3 of CTA 70; clocks relative to item 2's first stamp"}))
    rows = []
    for label, its in (("item2", it2), ("item3", it3)):
        for j, (e, m) in enumerate(its):
            rows.append({"item": label, "j": j, "grp": "?", "c_start": e[0] - t0, "c_wait_sdp": e[1] - e[0], "c_ld": e[2] - e[1], "c_math": e[3] - e[2],
                         "c_wait_acc": e[4] - e[3], "c_store": e[5] - e[4], "m_sdp_issued": m[0] - t0, "m_dvdk_issued": m[1] - t0})
    show = rows if len(rows) <= 24 else rows[:3] + rows[nb - 3:nb + 3] + rows[-2:]
    for r in show:
        print(json.dumps(r))
    for g in (0, 1):
        x = t[8 + g * 8:8 + g * 8 + 5]
        print(json.dumps({"transition_item2_group": g, "loop_left": x[0] - t0, "kv_free_seen": x[1] - t0, "parked": x[2] - t0, "all_done_seen": x[3] - t0,
                          "epilogue_stored": x[4] - t0}))
    if nb > 2:
        print(json.dumps({"item_period_clk": (it3[0][0][0] or 0) - (it2[0][0][0] or 0), "alt_period_from_mma": it3[0][1][0] - it2[0][1][0]}))


if __name__ == "__main__":
    main()
