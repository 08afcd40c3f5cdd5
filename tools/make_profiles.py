"""Summarise ncu captures under gpurun_out/ into tracked files under profiles/.

    python tools/make_profiles.py r01 name=report.ncu-rep[:workload:kernel] ...   launches=launches.csv

For each report: profiles/<round>_<name>.md (key metrics + top stall lines) and an entry in
profiles/traffic.json (dram bytes read+written per launch, used by bench.py's roofline.traffic).
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rnd = sys.argv[1]
    os.makedirs(OUT, exist_ok=True)
    tpath = os.path.join(OUT, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for arg in sys.argv[2:]:
        name, val = arg.split("=", 1)
        if name == "launches":
            rows = list(csv.reader(open(val)))
            h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
            per = {}
            for r in rows[h + 1:]:
                per.setdefault(r[4].split("(")[0], []).append(float(r[-1]))
            tot = sum(sum(v) for v in per.values())
            with open(os.path.join(OUT, f"{rnd}_launches.md"), "w") as f:
                f.write(f"# {rnd}: launch list of `python bench.py` (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n"
                        "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|\n")
                for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
                    f.write(f"| {k} | {len(v)} | {sum(v)/1e3:.1f} | {sum(v)/len(v)/1e3:.1f} | {100*sum(v)/tot:.1f}% |\n")
            with open(os.path.join(OUT, f"{rnd}_launches.csv"), "w") as f:
                f.write(open(val).read())
            continue
        parts = val.split(":")
        rep = parts[0]
        raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
        hdr, units, d = raw[0], raw[1], raw[2]
        m = dict(zip(hdr, d)); u = dict(zip(hdr, units))
        def num(k):
            return float(m[k].replace(",", "")) * UNIT.get(u.get(k, ""), 1.0)
        rd, wr, dur = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), float(m["gpu__time_duration.sum"].replace(",", ""))
        dur_s = dur * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u["gpu__time_duration.sum"], 1e-6)
        lines = [f"# {rnd}: ncu --set full --clock-control none — {m['Kernel Name']}", "",
                 f"grid {m['Grid Size']} block {m['Block Size']}; report `{os.path.basename(rep)}` (scratch, not tracked)", "",
                 f"- dram read+write per launch: {(rd + wr)/1e9:.4f} GB  ->  {(rd + wr)/dur_s/1e9:.0f} GB/s under the profiler (replayed, cold cache)", ""]
        lines += ["| metric | value | unit |", "|---|---|---|"] + [f"| {k} | {m.get(k)} | {u.get(k)} |" for k in KEYS if k in m]
        lines += ["", "stall reasons (warp-cycles per issued instruction):", ""]
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
                try:
                    v = float(m[k])
                except ValueError:
                    continue
                if v >= 0.2:
                    lines.append(f"- {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}: {v:.2f}")
        src = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"])
        rows = list(csv.reader(io.StringIO(src)))
        if len(rows) > 3:
            h2 = rows[2]; ix = {}
            for i, hh in enumerate(h2):
                ix.setdefault(hh, i)
            L = []
            for r in rows[3:]:
                if len(r) >= len(h2) and r[0]:
                    try:
                        L.append((int(r[0]), r[1].strip(), int(r[ix['# Samples']] or 0), int(r[ix['Instructions Executed']] or 0)))
                    except ValueError:
                        pass
            ti = sum(x[3] for x in L) or 1; ts = sum(x[2] for x in L) or 1
            lines += ["", f"source lines by stall samples (total warp-instructions {ti}):", "", "| line | % inst | % samples | source |", "|---|---|---|---|"]
            for x in sorted(L, key=lambda x: -x[2])[:12]:
                lines.append(f"| {x[0]} | {100*x[3]/ti:.1f} | {100*x[2]/ts:.1f} | `{x[1][:100].replace('|', '/')}` |")
        with open(os.path.join(OUT, f"{rnd}_{name}.md"), "w") as f:
            f.write("\n".join(lines) + "\n")
        if len(parts) == 3:
            traffic.setdefault(parts[1], {})[parts[2]] = rd + wr
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
