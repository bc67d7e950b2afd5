"""FP64 instructions per fine step of the evaluation kernels, counted in the SASS of the built library.

    python tools/sass_counts.py            -> profiles/r2_sass_fp64_counts.json + profiles/r2_sass_inner_loops.md

For every kernel listed below the innermost backward-branch loops are found (a backward BRA whose body contains no other
backward branch), the one(s) with the most FP64 instructions are the time loop (unrolled LNA_UNROLL = 2 fine steps); the
script reports DFMA / DMUL / DADD per fine step, the other instructions of the loop, the executed flops per step
(2 DFMA + DMUL + DADD) and the dispatch cycles per warp-step (2 per FP64 instruction -- 16 FP64 lanes per SM sub-partition
-- plus 1 per other instruction).  bench.py turns these counts into `roofline.pipe_util`.
"""
import json, os, re, subprocess, sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "lnaphylodyn_b200", "liblnaphylodyn_b200.so")
KERNELS = {
    "sir_full_nostore": "void lna::k_full_loglik<0, false>",
    "sir_full_store": "void lna::k_full_loglik<0, true>",
    "sir_full_balanced_nostore": "void lna::k_full_loglik_bal<false>",
    "seir_full_nostore": "void lna::k_full_loglik<1, false>",
    "sirs_full_nostore": "void lna::k_sirs_full<false>",
    "cand_ess_par_w8": "void lna::k_cand_eval<3, 8, false>",
    "cand_ess_par_w16": "void lna::k_cand_eval<3, 16, false>",
    "cand_mh_pair": "void lna::k_cand_eval<5, 4, false>",
}
UNROLL = 2


def sass_functions():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            funcs[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m and name:
            funcs[name].append((int(m.group(1), 16), m.group(2).strip()))
    return funcs


def opcode(ins):
    parts = ins.split()
    return parts[1] if parts[0].startswith("@") else parts[0]


def loops(ins):
    """innermost loops as (start, end) address pairs"""
    back = []
    for ad, text in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s+)?(?:P\d,\s+)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) <= ad:
            back.append((int(m.group(1), 16), ad))
    inner = [(a, b) for a, b in back if not any((a2, b2) != (a, b) and a <= a2 and b2 <= b for a2, b2 in back)]
    return inner


def main():
    funcs = sass_functions()
    out, md = {}, ["# Round 2 -- FP64 instructions per fine step, counted in the SASS (`tools/sass_counts.py`)", "",
                   "`cuobjdump -sass lnaphylodyn_b200/liblnaphylodyn_b200.so`, innermost time loop of each kernel (unrolled x2).",
                   "Dispatch cycles per warp-step = 2 x FP64 instructions (16 FP64 lanes per SM sub-partition) + other instructions.", ""]
    for key, prefix in KERNELS.items():
        name = next((n for n in funcs if n.startswith(prefix)), None)
        if name is None:
            continue
        ins = funcs[name]
        best = None
        for a, b in loops(ins):
            body = [t for ad, t in ins if a <= ad <= b]
            c = Counter(opcode(t).split(".")[0] for t in body)
            fp = c["DFMA"] + c["DMUL"] + c["DADD"]
            if best is None or fp > best[0]:
                best = (fp, c, a, b, body)
        fp, c, a, b, body = best
        other = sum(c.values()) - fp
        # fine steps per trip = the loop counter's increment (VIADD R, R, imm / IADD3 R, R, imm, RZ on the same register)
        unroll = UNROLL
        for t in body:
            m = re.match(r"(?:VIADD|IADD3?) R(\d+), R(\d+), (0x[0-9a-f]+|\d+)", t.split(";")[0])
            if m and m.group(1) == m.group(2):
                unroll = int(m.group(3), 0)
        rec = {"kernel": name.split("(")[0], "loop": [hex(a), hex(b)], "unroll": unroll, "fp64_per_step": fp / unroll,
               "dfma_per_step": c["DFMA"] / unroll, "dmul_per_step": c["DMUL"] / unroll, "dadd_per_step": c["DADD"] / unroll,
               "other_per_step": other / unroll, "executed_flops_per_step": (2 * c["DFMA"] + c["DMUL"] + c["DADD"]) / unroll,
               "dispatch_cycles_per_warp_step": (2 * fp + other) / unroll}
        out[key] = rec
        md += [f"## `{rec['kernel']}`  loop {hex(a)}..{hex(b)}", "",
               f"per fine step: **{rec['fp64_per_step']:g} FP64** (DFMA {rec['dfma_per_step']:g}, DMUL {rec['dmul_per_step']:g}, "
               f"DADD {rec['dadd_per_step']:g}) + {rec['other_per_step']:g} other; executed flops {rec['executed_flops_per_step']:g}; "
               f"dispatch cycles per warp-step {rec['dispatch_cycles_per_warp_step']:g}", "", "```"] + body + ["```", ""]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_sass_fp64_counts.json"), "w"), indent=1)
    open(os.path.join(ROOT, "profiles", "r2_sass_inner_loops.md"), "w").write("\n".join(md) + "\n")
    for k, v in out.items():
        print(f"{k:22s} FP64/step {v['fp64_per_step']:6.1f}  other {v['other_per_step']:5.1f}  flops {v['executed_flops_per_step']:6.1f}")


if __name__ == "__main__":
    main()
