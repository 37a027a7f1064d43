#!/usr/bin/env python
"""FP64-pipe instructions per loop body of the render kernel, counted in the SASS of the built library (no GPU).

    python tools/loop_body_fp64.py [--lib <libschw_b200.so>] [--out <json>]

`nvdisasm -gi` prints, for every instruction, the chain of inlined call sites it belongs to.  The instructions of one
loop body of `persistent_rays<RenderJob, true, true>` are those whose chain passes through the kernel's call of
`step_BCBA_planar` (schw_kernels.cu); inside it they are split by the call site in `step_BCBA_planar` (schw_device.cuh)
they were inlined from: the planar tails (`flow_tail_planar`, `mix_planar`) run only on the fast path, the general tails
(`flow_tail`, `memo_refresh`, the theta `mix`) only off it, everything else on both.  bench.py multiplies these counts by
the loop bodies the device counted (Counters::warp_iters / planar_iters) to get the FP64-pipe utilisation of a run; the
JSON is keyed by the SHA of csrc/ so a stale file is detected, and __graft_entry__.build() regenerates it.
"""
import argparse
import collections
import hashlib
import json
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200")
FP64 = ("DFMA", "DMUL", "DADD", "DSETP")
# (kernel name substring, call marker in schw_kernels.cu, rays per thread): the shipped render kernel first
KERNELS = (("persistent_tilesINS_9RenderJobELi2E", "step_BCBA_planar_n<R>(", 2),
           ("persistent_tilesINS_9RenderJobELi1E", "step_BCBA_planar_n<R>(", 1),
           ("persistent_raysINS_9RenderJobELb1ELb1", "step_BCBA_planar(", 1))


def csrc_sha():
    hsh = hashlib.sha256()
    for f in ("schw_kernels.cu", "schw_device.cuh", "Makefile"):
        hsh.update(open(os.path.join(PKG, "csrc", f), "rb").read())
    return hsh.hexdigest()[:16]


def analyse(txt, ksrc, dsrc, kernel, marker, rays):
    call_lines = [i + 1 for i, l in enumerate(ksrc) if marker in l and "//" not in l.split(marker)[0]]
    assert call_lines, f"no call of {marker} in schw_kernels.cu"
    inside = False
    chain = []
    fresh = True            # the next //## line starts a new chain
    cat_fp = collections.Counter(); cat_all = collections.Counter()
    bra_div = 0
    for line in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            inside = kernel in m.group(1)
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            if fresh:
                chain = []
                fresh = False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.search(r"^\s+/\*([0-9a-f]+)\*/\s+((?:@!?U?P\d\s+)?)([A-Z0-9_.]+)(.*?);", line)
        if not m:
            continue
        fresh = True
        bra_div += m.group(3).startswith("BRA.DIV")
        hits = [c for c in chain if c[0] == "schw_kernels.cu" and c[1] in call_lines]
        if not hits:
            continue
        # the call site inside the step function = the chain entry right below the kernel's call
        k = chain.index(hits[0])
        site = chain[k - 1] if k > 0 else None
        text = dsrc[site[1] - 1] if site and site[0] == "schw_device.cuh" else ""
        deeper = " ".join(dsrc[c[1] - 1] for c in chain[:k] if c[0] == "schw_device.cuh")
        if "trig_factors(" in deeper or "sincos_flow(" in deeper or "m.hcis = hh * cis" in deeper:
            cat = "memo_miss"          # re-evaluation of a memoised trig factor: only when an angle's bits changed
        elif "flow_tail_planar(" in text or "mix_planar(" in text:
            cat = "fast_only"
        elif ("flow_tail(" in text or "memo_refresh(" in text or re.search(r"if \(!v\d\) mix\(", text) or "++n_general" in text
              or re.search(r"mix\(S\[r\]\.q1h", text)):
            cat = "general_only"
        else:
            cat = "common"
        op = m.group(3).split(".")[0]
        cat_all[cat] += 1
        if op in FP64:
            cat_fp[cat] += 1
    assert cat_fp["common"] > 0, f"no instruction of the loop body of {kernel} found: did the kernel or the step function change its name?"
    return {
        "kernel": kernel, "rays_per_thread": rays,
        # per ray-step (a loop body of an R-ray kernel advances R rays; Counters::warp_iters counts ray slots)
        "fp64_per_body_fast_path": (cat_fp["common"] + cat_fp["fast_only"]) / rays,
        "fp64_per_body_general": (cat_fp["common"] + cat_fp["general_only"]) / rays,
        "fp64_per_memo_miss_all_three_flows": cat_fp["memo_miss"] / rays,
        # BRA.DIV = the compiler could not prove that a *_sync vote is reached convergently and emitted divergence handling
        # around it (a 13 % slowdown of the step loop when it happened); must stay 0 (tests/test_abi_symbols.py)
        "bra_div_in_kernel": bra_div,
        "detail": {"fp64": dict(cat_fp), "static_instructions": dict(cat_all)},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(PKG, "lib", "libschw_b200.so"))
    ap.add_argument("--out", default=os.path.join(PKG, "lib", "loop_body.json"))
    a = ap.parse_args()
    ksrc = open(os.path.join(PKG, "csrc", "schw_kernels.cu")).read().splitlines()
    dsrc = open(os.path.join(PKG, "csrc", "schw_device.cuh")).read().splitlines()
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(a.lib)], cwd=d, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], cwd=d, capture_output=True, text=True).stdout
    per_kernel = [analyse(txt, ksrc, dsrc, *k) for k in KERNELS]
    # BRA.DIV in every kernel of namespace schw (cub's sort/scan kernels have their own and are not ours)
    bra_div_all, cur = collections.Counter(), None
    for line in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            cur = m.group(1) if m.group(1).startswith("_ZN4schw") else None
            if cur:
                bra_div_all[cur] += 0
        elif cur and "BRA.DIV" in line:
            bra_div_all[cur] += 1
    out = dict(per_kernel[0])                 # the shipped render kernel at the top level (what bench.py reads)
    out.update({
        "csrc_sha": csrc_sha(), "kernels": per_kernel,
        "bra_div_all_schw_kernels": sum(bra_div_all.values()), "schw_kernels": len(bra_div_all),
        "how": "nvdisasm -gi of the built library: instructions whose inline chain passes through the kernel's call of the step "
               "function (step_BCBA_planar / step_BCBA_planar_n<R>), split by call site (tools/loop_body_fp64.py), per ray-step.  "
               "'general' = all three flows and the mixing off the fast path with the trig memo hitting; the memo-miss "
               "re-evaluation (sincos + reciprocal, only when an angle's bits changed) is listed separately and not counted by "
               "bench.py.  The loop control outside the step function (termination pre-test, counters) holds no FP64-pipe instruction",
    })
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
