"""Turn the raw files of an evidence run (scripts/r2_evidence.sh, brought back in gpurun_out/) into the committed summaries under
profiles/: copies the bench lines / logs, writes the per-kernel share table of the ncu launch list, the per-launch conv metric table,
the `--set full` summary + per-source-line stall attribution of the dominant kernel, and profiles/<tag>_conv_dram_traffic.json
(DRAM bytes per launch of every conv class, keyed to the digest of the kernel sources so bench.py only reports it for the same code).

    python scripts/summarize_evidence.py [tag=r02]
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
        "byte/s": 1e-12, "Kbyte/s": 1e-9, "Mbyte/s": 1e-6, "Gbyte/s": 1e-3, "Tbyte/s": 1.0}


def copy(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, f"{TAG}_{dst}"))


def metric_rows(path):
    rows = list(csv.reader(open(path)))
    i0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[i0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[i0 + 1:]:
        if len(r) < len(hdr):
            continue
        k = int(r[ix["ID"]])
        per.setdefault(k, {"name": r[ix["Kernel Name"]]})
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            v = float("nan")
        per[k][r[ix["Metric Name"]]] = v * UNIT.get(r[ix["Metric Unit"]], 1.0)
    return per


def short(name):
    n = name.replace("void ", "").split("(")[0]
    return n.replace("tc5::", "").replace("tc4::", "").replace("tc3::", "").replace("(int)", "")


def launch_list():
    path = os.path.join(G, "ncu_launch_list_step.csv")
    if not os.path.exists(path):
        return
    shutil.copy(path, os.path.join(P, f"{TAG}_ncu_launch_list_step.csv"))
    per = metric_rows(path)
    by = collections.OrderedDict()
    for v in per.values():
        e = by.setdefault(short(v["name"]), [0, 0.0])
        e[0] += 1
        e[1] += v.get("gpu__time_duration.sum", 0.0)
    tot = sum(e[1] for e in by.values())
    with open(os.path.join(P, f"{TAG}_ncu_launch_list_step_summary.txt"), "w") as f:
        f.write(f"# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none: the launches of ONE timed denoising step\n"
                f"# (schedule position 0, LB2_GRAPHS=0 so every kernel is its own launch); {len(per)} launches, {tot / 1e3:.2f} ms summed (serialised, cold caches)\n")
        for k, (n, us) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * us / tot:6.2f} %  {us / 1e3:8.3f} ms  {n:4d} launches  {k}\n")


def conv_metrics():
    path = os.path.join(G, "ncu_conv_launch_metrics.csv")
    if not os.path.exists(path):
        return
    shutil.copy(path, os.path.join(P, f"{TAG}_ncu_conv_launch_metrics.csv"))
    per = metric_rows(path)
    cls = collections.defaultdict(lambda: [0, 0.0, 0.0])
    with open(os.path.join(P, f"{TAG}_ncu_conv_launch_metrics.txt"), "w") as f:
        f.write("# every sparse-convolution launch of one denoising step (schedule position 0), in launch order\n"
                "# (tensor-pipe activity is a --set full metric: see <tag>_ncu_pair_l3_full_summary.txt; n/a in this metrics-only pass)\n"
                "#  id kernel                          time[us]  L2->SM[TB/s]  DRAM read[MB]  write[MB]  L2 hit%\n")
        for k, v in per.items():
            tens = [x for kk, x in v.items() if "pipe_tensor_cycles_active" in kk]
            dr, dw = v.get("dram__bytes_read.sum", 0) / 1e6, v.get("dram__bytes_write.sum", 0) / 1e6
            f.write(f"{k:4d} {short(v['name']):30s} {v.get('gpu__time_duration.sum', 0):9.1f}  "
                    f"{v.get('l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 0):11.2f}  {dr:12.1f}  {dw:9.1f}  {v.get('lts__t_sector_hit_rate.pct', 0):6.1f}\n")
            c = cls[short(v["name"])]
            c[0] += 1
            c[1] += dr + dw
            c[2] += v.get("gpu__time_duration.sum", 0)
    import bench
    key = {"k_spconv_tc_pair<256>": "cout256", "k_spconv_tc_pair<128>": "cout128", "k_spconv_tc_n256": "cout256"}
    out = {"csrc_digest": bench.csrc_digest(), "source": f"profiles/{TAG}_ncu_conv_launch_metrics.csv (dram__bytes_read.sum + dram__bytes_write.sum, mean per launch)",
           "traffic_bytes_per_launch": {}, "per_kernel": {}}
    small = [0, 0.0]
    for name, (n, mb, us) in cls.items():
        out["per_kernel"][name] = {"launches": n, "mean_dram_bytes": mb * 1e6 / n, "mean_us": us / n}
        if name in key:
            out["traffic_bytes_per_launch"][key[name]] = mb * 1e6 / n
        elif "tc_small" in name:
            small[0] += n
            small[1] += mb
    if small[0]:
        out["traffic_bytes_per_launch"]["cout_le96"] = small[1] * 1e6 / small[0]
    json.dump(out, open(os.path.join(P, f"{TAG}_conv_dram_traffic.json"), "w"), indent=1)


def full_capture():
    rep = os.path.join(G, "prof_pair_l3.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{TAG}_ncu_pair_l3_full_raw.csv"), "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "launch__cluster_size", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
    with open(os.path.join(P, f"{TAG}_ncu_pair_l3_full_summary.txt"), "w") as f:
        f.write("# ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc_pair -s 12 -c 3 (LB2_GRAPHS=0)\n"
                "# launches = up1.1.0.net.0 (384->256, level 3), its 1x1 downsample (384->256), up1.1.0.net.3 (256->256) of the timed step (position 0)\n")
        for r in rows[2:]:
            f.write("-----\n")
            for w in ["Kernel Name", "Grid Size", "Block Size"] + want:
                for i, h in enumerate(hdr):
                    if h == w or h.endswith(w):
                        f.write(f"  {h} [{units[i]}] = {r[i]}\n")
                        break
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    tmp = os.path.join(G, "_src_tmp")
    os.makedirs(tmp, exist_ok=True)
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "lidiff_b200/_C/liblidiff_b200.so")], cwd=tmp, capture_output=True)
    sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, "spconv_tc5.sm_100a.cubin")], capture_output=True, text=True).stdout
    open(os.path.join(tmp, "tc5.sass"), "w").write(sass)
    with open(os.path.join(P, f"{TAG}_ncu_pair_l3_stall_attribution.txt"), "w") as f:
        f.write("# warp-stall samples per CUDA source line (scripts/ncu_src_lines.py: ncu --page source joined with nvdisasm -gi line tables)\n")
        for n, s in enumerate(starts):
            e = starts[n + 1] if n + 1 < len(starts) else len(rows)
            one = os.path.join(tmp, f"k{n}.csv")
            csv.writer(open(one, "w", newline="")).writerows(rows[s:e])
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts/ncu_src_lines.py"), one, os.path.join(tmp, "tc5.sass"), "k_spconv_tc_pairILi256", "24"],
                               capture_output=True, text=True)
            f.write(f"\n## launch {n}: {rows[s][1]}\n" + "\n".join(l[:200] for l in r.stdout.splitlines()) + "\n")
    shutil.rmtree(tmp, ignore_errors=True)


def main():
    for src, dst in (("bench_n1.json", "bench_n1.json"), ("bench_n1.stderr.log", "bench_n1.stderr.log"), ("bench_n1_steps50.json", "bench_n1_steps50.json"),
                     ("bench_reference_arm.json", "bench_reference_arm.json"), ("bench_T1000.json", "bench_T1000.json"), ("pytest_gpu.log", "pytest_gpu.log"),
                     ("smoke.log", "smoke.log"), ("env.txt", "env.txt")):
        copy(src, dst)
    launch_list()
    conv_metrics()
    full_capture()
    print("\n".join(sorted(f for f in os.listdir(P) if f.startswith(TAG))))


if __name__ == "__main__":
    main()
