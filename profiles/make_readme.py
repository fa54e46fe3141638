"""Builds profiles/README.md and the per-round summary files from the captures in gpurun_out/.

    python profiles/make_readme.py            # reads gpurun_out/{launches_r01_*_step.csv, ode_*.ncu-rep, sa_mlp2.ncu-rep, bench_*.json}

Copies the launch lists and the bench JSON lines into profiles/ (tracked); the .ncu-rep files stay in gpurun_out/
(scratch) and only their raw metrics / top source lines are summarised here.
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")
ROUND = "r02"


def launch_table(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg, total, n = collections.OrderedDict(), 0.0, 0
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki]))
        name = re.sub(r"<unnamed>::", "", name)[:70]
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    ours = sum(t for k, (c, t) in agg.items() if k.startswith(("gp::", "saf::", "gemm::")))
    out = [f"total {total / 1e3:.2f} ms in {n} launches; kernels of this library: {100 * ours / total:.1f} % of the time\n",
           "| kernel | launches | us | share |", "|---|---|---|---|"]
    other_c, other_t = 0, 0.0
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if t / total >= 0.004:
            out.append(f"| `{k}` | {c} | {t:.0f} | {100 * t / total:.1f}% |")
        else:
            other_c += c
            other_t += t
    out.append(f"| (everything below 0.4 %) | {other_c} | {other_t:.0f} | {100 * other_t / total:.1f}% |")
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]


def raw_metrics(rep, want=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"name": r[H.index("Kernel Name")]}
        for w in (want or WANT):
            if w in H:
                d[w] = f"{r[H.index(w)]} {U[H.index(w)]}".strip()
        res.append(d)
    return res


def top_lines(rep, n=14):
    out = subprocess.run([sys.executable, os.path.join(PR, "top_lines.py"), rep, str(n)], capture_output=True, text=True).stdout
    return "\n".join(l[:200] for l in out.splitlines())


def main():
    md = [f"# profiles — round 2\n",
          "All captures: one B200, `profiles/profile_step.py` (C2 workload: 64 objects x 50 hypotheses, T0 = 0.55), after 3 "
          "warm-up steps, one step between cudaProfilerStart/Stop.  Nsight Compute cannot replay a launch that is both "
          "cooperative and clustered, so the captures run with `GP_NONCOOPERATIVE_LAUNCH=1` (same grid, same kernels; "
          "DESIGN.md section 5).  `refresh.sh {core,solo,extra}` produces the captures on the GPU box, `make_readme.py` regenerates this "
          "file from them.  `sass_summary.txt` (from `sass_summary.py`): per-kernel counts of UTCHMMA / LDTM / STTM / UBLKCP / SYNCS in "
          "the shipped `.so`.\n"]
    for mode in ("fp32", "bf16"):
        src = os.path.join(GO, f"launches_{ROUND}_{mode}_step.csv")
        if os.path.exists(src):
            shutil.copy(src, os.path.join(PR, os.path.basename(src)))
            md.append(f"## launches_{ROUND}_{mode}_step.csv — every launch of one {mode}-mode step\n")
            md.append("`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python "
                      f"profiles/profile_step.py --mlp_mode {mode}` (cold-cache, serialised, the energy encoder does not overlap "
                      "the sampler under ncu: compare SHARES, not absolutes).\n")
            md.append(launch_table(src) + "\n")
    md.append("## ode_{fp32,bf16} — `ncu --set full` of the integrator kernel (`ode_rk45_kernel<TcEval<3>>`, `<TcEval<1>>`)\n")
    md.append("`ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ode_rk45 -c 1 python "
              "profiles/profile_step.py --what sampler --mlp_mode {fp32,bf16}`\n")
    reps = {m: os.path.join(GO, f"ode_{m}.ncu-rep") for m in ("fp32", "bf16")}
    mets = {m: raw_metrics(p)[0] for m, p in reps.items() if os.path.exists(p)}
    if mets:
        md.append("| metric | " + " | ".join(f"{m} (`{mets[m]['name'][:40]}`)" for m in mets) + " |")
        md.append("|---|" + "---|" * len(mets))
        for w in WANT:
            md.append(f"| {w} | " + " | ".join(mets[m].get(w, "") for m in mets) + " |")
        md.append("\nDRAM traffic is ~2 MB per launch (weights and state are L2 resident): the kernel is tensor / latency bound, not "
                  "HBM bound (`roofline.traffic` = dram read + write).  Algorithmic work per launch: 3200 rows x 154 evaluations x "
                  "0.5335 MFLOP = 0.26 TFLOP (issued three times in fp32 mode).  25 tiles x 4-CTA clusters = 100 of the 148 SMs.\n")
        for m, p in reps.items():
            if os.path.exists(p):
                md.append(f"### top source lines by warp-stall samples — {m}\n\n```\n{top_lines(p)}\n```\n")
                md.append("(`sm_20_intrinsics.hpp:151` / `tc_ptx.cuh:30-31` are the mbarrier wait loops: warps of one role waiting for "
                          "another role -- the chain gather/convert -> MMA -> epilogue of one evaluation is serial by data dependence.)\n")
    solo = {m: os.path.join(GO, f"ode_solo_{m}.ncu-rep") for m in ("fp32", "bf16")}
    smets = {m: raw_metrics(p)[0] for m, p in solo.items() if os.path.exists(p)}
    if smets:
        md.append("## ode_solo_{fp32,bf16} — `ncu --set full` of the one-CTA-per-tile integrator (`ode_rk45_kernel<TcSoloT<3>>`, `<TcSoloT<1>>`), "
                  "1024 objects x 50 hypotheses = 400 tiles (the per-GPU share of C5)\n")
        md.append("`ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ode_rk45 -c 1 python "
                  "profiles/profile_step.py --what sampler --objects 1024 --mlp_mode {fp32,bf16}`\n")
        md.append("| metric | " + " | ".join(f"{m} (`{smets[m]['name'][:44]}`)" for m in smets) + " |")
        md.append("|---|" + "---|" * len(smets))
        for w in WANT:
            md.append(f"| {w} | " + " | ".join(smets[m].get(w, "") for m in smets) + " |")
        md.append("\nAlgorithmic work per launch: 51 200 rows x 153 evaluations x 0.5335 MFLOP = 4.18 TFLOP (issued three times in fp32 "
                  "mode).  The activations (A operand) live in tensor memory, the weights stream through an 8-deep ring; DRAM traffic "
                  "is the float64 state (L2 resident after the first touch).\n")
        for m, p in solo.items():
            if os.path.exists(p):
                md.append(f"### top source lines by warp-stall samples — solo {m}\n\n```\n{top_lines(p)}\n```\n")
        md.append("(`trunk.cu:489` is the wait at the grid barrier: 104 CTAs own 3 tiles, 44 own 2.  In bf16 mode the head epilogue -- "
                  "`trunk_solo_t.cuh:350-361`, short-scoreboard stalls on the shared-memory loads of the output-layer weights -- is the "
                  "longer leg; in fp32 mode the epilogue warps wait for the MMA chains.)\n")
    c5 = os.path.join(GO, f"launches_{ROUND}_fp32_c5shard.csv")
    if os.path.exists(c5):
        shutil.copy(c5, os.path.join(PR, os.path.basename(c5)))
        md.append(f"## launches_{ROUND}_fp32_c5shard.csv — every launch of one fp32-mode step on a 1024-object shard\n")
        md.append("`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python profiles/profile_step.py "
                  "--objects 1024 --mlp_mode fp32`\n")
        md.append(launch_table(c5) + "\n")
    sa = os.path.join(GO, "sa_mlp2.ncu-rep")
    if os.path.exists(sa):
        md.append("## sa_mlp2 — `ncu --set full` of the fused set-abstraction kernel, the seven launches of one encoder\n")
        md.append("`ncu --set full -k regex:sa_mlp2 -c 7 python profiles/profile_step.py` (the wide scale of level 1 in first-level mode, then levels 2, 3, 4 x 2 scales; fp32 mode)\n")
        ms = raw_metrics(sa)
        keys = ["gpu__time_duration.sum", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
        md.append("| launch | " + " | ".join(k.split(".")[0] for k in keys) + " |")
        md.append("|---|" + "---|" * len(keys))
        for i, d in enumerate(ms):
            md.append(f"| {i} | " + " | ".join(d.get(k, "") for k in keys) + " |")
        md.append(f"\n### top source lines — all seven launches\n\n```\n{top_lines(sa)}\n```\n")
    sas = os.path.join(GO, "sa_small.ncu-rep")
    if os.path.exists(sas):
        md.append("## sa_small — `ncu --set full` of the level-1 kernel (weights as uniform operands from the launch's parameter space)\n")
        md.append("`ncu --set full -k regex:sa_small -c 2 python profiles/profile_step.py` (the two scales of one encoder)\n")
        keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
                "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum"]
        ms = raw_metrics(sas, keys)
        md.append("| kernel | " + " | ".join(k.split(".")[0] for k in keys) + " |")
        md.append("|---|" + "---|" * len(keys))
        for d in ms:
            md.append(f"| `{re.sub(r'[(].*', '', d['name'])[5:]}` | " + " | ".join(d.get(k, "") for k in keys) + " |")
        md.append("\nThe wide scale issues on 83 % of the cycles and the FMA pipe is busy 59 % of them: it is issue-bound, the "
                  "rest of the slots are the `LDCU.128` that feed the uniform registers (one per four FFMAs), the pooling `REDUX` "
                  "and the gather.  With the weights staged in shared memory (`gp_sa_small_mlp`) the same launch took 267 us.\n")
    geo, geolog = os.path.join(GO, "geom.ncu-rep"), os.path.join(GO, "geom.log")
    if os.path.exists(geo) and os.path.exists(geolog):
        md.append("## geom — `ncu --set full` of FPS / ball query / grouping, the four encoder levels at 64 objects\n")
        md.append("`ncu --profile-from-start off --set full --clock-control none -k regex:'fps_kernel|ball_query|group_kernel' "
                  "python profiles/profile_geometry.py`.  `alg. bytes` are the algorithmic bytes of DESIGN.md 4.1-4.3 (cloud + centres "
                  "read, indices / grouped tensor written); `GB/s` = alg. bytes / duration, against the measured copy bandwidth in "
                  f"`MEASURED_PEAKS.json`; `dram` = dram__bytes_read + dram__bytes_write of the same launch (cold cache).\n")
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        ops = [json.loads(l) for l in open(geolog) if l.startswith("{")]
        ms = raw_metrics(geo)
        md.append("| op | level | kernel | grid | us | alg. bytes | GB/s | % of HBM peak | dram bytes | sm throughput % | warps active % |")
        md.append("|---|---|---|---|---|---|---|---|---|---|---|")

        def num(d, k):
            v, u = (d.get(k, "0 ") + " ").split(" ")[:2]
            v = float(v.replace(",", ""))
            return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
        for o, d in zip(ops, ms):
            us = num(d, "gpu__time_duration.sum")
            dram = num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum")
            gbs = o["algorithmic_bytes"] / us / 1e3
            md.append(f"| {o['op']} | {o['level'] + 1} | `{re.sub(r'[(].*', '', d['name'])[:34]}` | {d.get('launch__grid_size', '').split(' ')[0]} | {us:.1f} | "
                      f"{o['algorithmic_bytes']:,} | {gbs:.0f} | {100 * gbs / peak:.1f} | {dram:,.0f} | "
                      f"{num(d, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
                      f"{num(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} |")
        md.append("\nFPS is a chain of npoint - 1 dependent arg-max steps per object on 64 CTAs (one per object): it is bound by the "
                  "latency of one step (distance update, two redux levels, one barrier), not by bytes -- 0.6 MB per launch; "
                  "`fps_chain` is what the encoder calls (levels 2-4 take the FPS-order prefix, DESIGN.md 4.1). "
                  "Ball query is issue-bound (one thread per centre, 11.7 instructions per centre-point pair); grouping at these "
                  "sizes reaches 53-63 % of the copy bandwidth, and the encoder's hot path no longer materialises the grouped "
                  "tensor at all (DESIGN.md 4.4).\n")
    def last_json(path):
        return json.loads([l for l in open(path) if l.startswith("{")][-1])
    for name in ("bench_fp32.json", "bench_ffma.json", "bench_reference.json", "bench_2gpu.json", "bench_4gpu.json", "bench_8gpu.json"):
        src = os.path.join(GO, name)
        if os.path.exists(src):
            with open(os.path.join(PR, f"bench_{ROUND}_{name[6:]}"), "w") as f:
                f.write(json.dumps(last_json(src)) + "\n")
    b = os.path.join(GO, "bench_fp32.json")
    if os.path.exists(b):
        j = last_json(b)
        md.append(f"## bench_{ROUND}_*.json — `python bench.py --steps 20 --warmup 3`, `--impl reference`, `--mlp_mode fp32_ffma`, `torchrun ... --gpus N`\n")
        rf = j["roofline"]
        md.append(f"fp32 mode: {j['value']:.0f} {j['unit']} ({j['ms_per_step']:.2f} ms/step, {j['config'].get('cuda_graph', '')}), e2e {j['e2e']['value']:.0f}; "
                  f"integrator kernel {rf['kernel_ms']:.2f} ms = {rf['achieved']:.1f} TFLOP/s algorithmic "
                  f"({100 * rf['frac']:.1f} % of the burst bf16 tensor peak, {100 * rf.get('frac_of_sustained', 0):.1f} % of the sustained one, x3 issued); bf16 mode: "
                  f"{j['other_mode']['value']:.0f} {j['unit']} ({j['other_mode']['ms_per_step']:.2f} ms/step); cpu_baseline ({j['cpu_baseline']['kind']}): "
                  f"{j['cpu_baseline']['value']:.1f} {j['unit']} on {j['cpu_baseline']['cores']} cores; clocks {j['clocks']}.\n")
        oc = j.get("other_configs", {})
        if "c5" in oc and "fp32" in oc["c5"]:
            c = oc["c5"]
            md.append(f"C5 on one GPU ({c['objects_total']} objects x 50, strong-scaling point N = 1): fp32 {c['fp32']['value']:.0f} objects/s "
                      f"({c['fp32']['ms_per_step']:.0f} ms/step; integrator {c['fp32']['roofline']['achieved']:.0f} TFLOP/s algorithmic = "
                      f"{100 * c['fp32']['roofline']['frac']:.1f} % of the burst peak, x3 issued), bf16 {c['bf16']['value']:.0f} "
                      f"({c['bf16']['ms_per_step']:.0f} ms/step; {c['bf16']['roofline']['achieved']:.0f} TFLOP/s = {100 * c['bf16']['roofline']['frac']:.1f} %).\n")
        if "c1" in oc and "fp32" in oc["c1"]:
            md.append(f"C1 (1 object, sampler only, T0 = 1): {oc['c1']['fp32']['latency_ms']:.2f} ms; reference cond_ode_sampler: "
                      f"{oc['c1'].get('reference_cpu_ms', 0):.0f} ms on the CPU, {oc['c1'].get('reference_gpu_ms', 0):.0f} ms on this GPU.  "
                      f"C4 (tracking, 32 objects/frame, 100 frames): {oc.get('c4', {}).get('ms_per_frame', 0):.2f} ms/frame.\n")
        if "c3" in oc and "sweep" in oc["c3"]:
            md.append("C3 sweep (256 objects; ours vs the reference's own CUDA ext on the same clouds, us; indices bit-exact):\n")
            md.append("| points | FPS | ref FPS | ball query x2 | ref | query+group C=3 | ref group | group C=96 | ref |")
            md.append("|---|---|---|---|---|---|---|---|---|")
            for r in oc["c3"]["sweep"]:
                g96 = r.get("group_C96", {})
                md.append(f"| {r['points']} | {r['fps']['us']:.0f} | {r['fps'].get('reference_ext_us', 0):.0f} | {r['ball_query_2radii']['us']:.0f} | "
                          f"{r['ball_query_2radii'].get('reference_ext_us', 0):.0f} | {r['query_group_C3']['us']:.0f} | "
                          f"{r['query_group_C3'].get('reference_ext_group_only_us', 0):.0f} | {g96.get('us', 0):.0f} | {g96.get('reference_ext_us', 0):.0f} |")
            md.append("")
        rg = j.get("reference_gpu") or {}
        if "c2_full_path" in rg:
            md.append(f"Reference on this GPU (unmodified modules + its CUDA ext, torch fp32 + host scipy loop + sklearn): full path "
                      f"{rg['c2_full_path']['objects_per_s']:.0f} objects/s ({rg['c2_full_path']['ms_per_step']:.0f} ms/step, stages "
                      f"{ {k: round(v) for k, v in rg['c2_full_path']['stages_ms'].items()} }), sampling {rg['c2_full_path']['sampling_objects_per_s']:.0f} objects/s; "
                      f"this library e2e / reference GPU full path = {rg.get('speedup_e2e_vs_reference_gpu_full_path', 0):.0f} x.\n")
        br = os.path.join(GO, "bench_reference.json")
        if os.path.exists(br):
            r = last_json(br)
            md.append(f"`--impl reference` ({r['cpu_baseline']['kind']}): {r['value']:.2f} objects/s ({r['ms_per_step'] / 1e3:.2f} s/step, "
                      f"{r['steps_timed']} steps, {r['cpu_baseline']['cores']} threads): {r['cpu_baseline']['sample']}.\n")
        for n in (2, 4, 8):
            bn = os.path.join(GO, f"bench_{n}gpu.json")
            if os.path.exists(bn):
                jn = last_json(bn)
                c = jn.get("other_configs", {}).get("c5", {})
                extra = ""
                if "fp32" in c:
                    extra = (f"; C5 strong scaling (8192 objects over {n} ranks, gather in the timed region): fp32 {c['fp32']['value']:.0f} objects/s "
                             f"({c['fp32']['ms_per_step']:.0f} ms/step), bf16 {c['bf16']['value']:.0f}")
                md.append(f"bench_{ROUND}_{n}gpu.json (`torchrun --nproc-per-node {n} bench.py --gpus {n}`): weak C2 {jn['value']:.0f} {jn['unit']} = "
                          f"{jn['value'] / j['value']:.2f} x the one-GPU line ({jn['ms_per_step']:.2f} ms/step, max over ranks){extra}.\n")
    md.append("## phase_breakdown.py — in-kernel cycle counters of the integrator\n")
    md.append("`python profiles/phase_breakdown.py` reads `stats[8..24]` of `gp_scorenet_ode` (cycles per RHS evaluation, CTA 0 of "
              "cluster 0; forward = l1 + wait_d1 + epi1 + wait_heads + epi2 + the x:* exchange tail; the stage_* / err rows are the "
              "integrator between evaluations).\n")
    ph = os.path.join(GO, "phase.log")
    if os.path.exists(ph):
        shutil.copy(ph, os.path.join(PR, f"phase_breakdown_{ROUND}.txt"))
        md.append("```\n" + open(ph).read().strip() + "\n```\n")
    open(os.path.join(PR, "README.md"), "w").write("\n".join(md))
    print("wrote profiles/README.md")


if __name__ == "__main__":
    main()
