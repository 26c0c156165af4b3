"""Builds profiles/r2_traffic.json (DRAM bytes per launch of the dominant stages, read by bench.py into roofline.traffic)
from `ncu --set full` reports of tools/profile_solve.py:

    python tools/make_traffic_json.py cube118=gpurun_out/r2_prof_full_118.ncu-rep cube255=gpurun_out/r2_prof_full_255.ncu-rep

A smoothing stage of the fine level = its size-class launches (consecutive smooth_ell_kernel launches with different
block sizes); inside a V-cycle the first group is the pre-smoothing stage, the second the post-smoothing stage."""
import csv, json, os, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik, ib = hdr.index("Kernel Name"), hdr.index("Block Size")
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    res = []
    for r in rows[2:]:
        b = float(r[ir].replace(",", "")) * UNIT.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * UNIT.get(units[iw], 1.0)
        res.append((r[ik], r[ib], b, float(r[it].replace(",", ""))))
    return res


def stages(ls):
    """pre stage = a run of smooth_ell_kernel launches followed by the A_out update of the residual (sell_spmv_kernel<3, ..>);
    post stage = any other complete run (runs cut by the capture window — fewer launches than the longest run — are dropped)."""
    runs, cur = [], []
    for k, (name, blk, b, t) in enumerate(ls):
        if "smooth_ell_kernel" in name:
            cur.append(b)
        elif cur:
            runs.append((cur, name)); cur = []
    full = max((len(r) for r, _ in runs), default=0)
    is_res_out = lambda nm: "sell_spmv_kernel<3" in nm or "sell_spmv_kernel<(int)3" in nm
    pre = [sum(r) for r, nxt in runs if len(r) == full and is_res_out(nxt)]
    post = [sum(r) for r, nxt in runs if len(r) == full and not is_res_out(nxt)]
    dot = [b for name, blk, b, t in ls if "sell_spmv_kernel<0, 1>" in name or "sell_spmv_kernel<(int)0, (bool)1>" in name]
    mean = lambda v: int(sum(v) / len(v)) if v else None
    return {"pre_smooth@level0": mean(pre), "post_smooth@level0": mean(post), "spmv_dot@level0": mean(dot),
            "stages_captured": {"pre": len(pre), "post": len(post), "spmv_dot": len(dot)}}


out = {"source": "ncu --set full --clock-control none of tools/profile_solve.py (dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches; "
                 "a smoothing stage = its size-class launches)"}
for arg in sys.argv[1:]:
    key, rep = arg.split("=")
    out[key] = stages(launches(rep))
    out[key]["report"] = os.path.basename(rep)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
