"""Prints the essentials of bench lines: python tools/show_bench.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        p = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    c = p["config"]
    print(f"== {f}: {p['n_gpus']} GPU(s)  {p['ms_per_step']:.2f} ms/solve  {p['value']/1e6:.1f} M DOF/s  iters {c['iterations']}  single {c.get('single_gpu_same_mesh_ms')}  "
          f"speedup {c.get('speedup_vs_single_gpu_same_mesh')}  e2e {p['e2e']['ms_per_step']:.2f} ms")
    print("   setup", round(c["setup_ms"], 1), "pattern", round(c["pattern_ms"], 1), "assemble", round(c["assemble_ms"], 1), " incl-setup", round(c["dofs_per_s_incl_assembly_and_setup"] / 1e6, 1), "M DOF/s")
    print("   parity", json.dumps(p.get("parity"))[:400])
    r = p["roofline"]
    print("   roofline", r["kernel"], round(r["frac"], 3), "us", round(r["us_per_launch"], 1))
    for k, v in r["top_kernels"].items():
        print("     ", k, v)
    if r.get("us_per_launch_min_max_over_ranks"):
        print("   min/max over ranks (us):", r["us_per_launch_min_max_over_ranks"])
    if p.get("secondary"):
        s = p["secondary"]
        print("   secondary", s["ms_per_step"], s["roofline"]["kernel"], round(s["roofline"]["frac"], 3), "setup", s["setup_ms"], "scipy", s.get("scipy_cg_jacobi"))
