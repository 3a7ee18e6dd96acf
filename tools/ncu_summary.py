"""Summarise ncu CSV exports into the small tables committed under profiles/.

  python tools/ncu_summary.py launches <ncu --csv launch list> <out.csv>      per-kernel avg gpu__time_duration + share
  python tools/ncu_summary.py full <ncu -i rep --page raw --csv> <out.csv> [<inst_counts.json> <key>]
"""
import csv, json, sys, re, collections

mode, src, out = sys.argv[1:4]
rows = [r for r in csv.reader(open(src, errors="ignore")) if r]
if mode == "launches":
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value"); mn = hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[hdr_i + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("hdgnn::", "")
        a = agg.setdefault(name, [0.0, 0]); a[0] += float(r[mv].replace(",", "")); a[1] += 1
    unit_ns = True
    tot = sum(a[0] for a in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,avg_us,share\n")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"\"{n}\",{a[1]},{a[0] / a[1] / 1e3:.2f},{a[0] / tot:.3f}\n")
    print(open(out).read())
else:
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print(open(out).read())
    if len(sys.argv) > 5:
        jpath, key = sys.argv[4:6]
        try:
            ic = json.load(open(jpath))
        except Exception:
            ic = {}
        label = {"ent_fwd2_kernel": "ent_fwd", "ent_bwd2_kernel": "ent_bwd", "mid2_kernel": "mid(train)", "pack_bits_kernel": "pack_bits",
                 "reduce_adam_kernel": "reduce_adam"}
        kn = hdr.index("Kernel Name"); ii = hdr.index("smsp__inst_executed.sum")
        dr, dw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        d = ic.setdefault(key, {})
        for r in rows[2:]:
            for k, lab in label.items():
                if k in r[kn]:
                    d[lab] = {"inst": float(r[ii].replace(",", "")),
                              "dram_bytes": float(r[dr].replace(",", "")) * scale.get(units[dr], 1) + float(r[dw].replace(",", "")) * scale.get(units[dw], 1)}
        ic["_source"] = "ncu --set full --clock-control none: smsp__inst_executed.sum, dram__bytes_read.sum + dram__bytes_write.sum per launch"
        json.dump(ic, open(jpath, "w"), indent=1)
