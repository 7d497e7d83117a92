"""Summarise ncu output for profiles/: (1) a per-launch metrics CSV of one reverse step mapped onto the op list that
bench.py --dump-ops wrote (launch order == op order), (2) the raw page of a --set full capture.

    python tools/ncu_summary.py step <metrics.csv> <ops.json> <out.json>
    python tools/ncu_summary.py full <file.ncu-rep> <out.json>
"""
import csv
import json
import subprocess
import sys

OURS = ("conv_tma", "conv_ws", "conv_ffma", "attention_", "head_kernel", "encode_input")


def step(metrics_csv, ops_json, out_json):
    rows = [r for r in csv.reader(open(metrics_csv)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rows[1:]:
        d = launches.setdefault(r[ix["ID"]], dict(kernel=r[ix["Kernel Name"]], grid=r[ix["Grid Size"]], block=r[ix["Block Size"]]))
        try:
            d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            d[r[ix["Metric Name"]]] = None
        d[r[ix["Metric Name"]] + ".unit"] = r[ix["Metric Unit"]]
    seq = [launches[k] for k in sorted(launches, key=int)]
    ops = json.load(open(ops_json))
    n = len(ops)
    # the first reverse step starts right after the first time_table launch
    start = next(i for i, l in enumerate(seq) if "time_table" in l["kernel"]) + 1
    seq = [l for l in seq[start:] if any(k in l["kernel"] for k in OURS)][:n]
    assert len(seq) == n, (len(seq), n)
    out, agg = [], {}
    total_ns = sum(l["gpu__time_duration.sum"] for l in seq)
    for o, l in zip(ops, seq):
        dram = (l.get("dram__bytes_read.sum") or 0) + (l.get("dram__bytes_write.sum") or 0)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = (l.get("dram__bytes_read.sum") or 0) * scale.get(l.get("dram__bytes_read.sum.unit", "byte"), 1) + \
               (l.get("dram__bytes_write.sum") or 0) * scale.get(l.get("dram__bytes_write.sum.unit", "byte"), 1)
        t = l["gpu__time_duration.sum"] * {"ns": 1, "us": 1e3, "ms": 1e6}.get(l.get("gpu__time_duration.sum.unit", "ns"), 1)
        rec = dict(index=o["index"], op_class=o["op_class"], kernel=l["kernel"].split("(")[0][-40:], grid=l["grid"], ns=t, dram_bytes=dram,
                   algorithmic_bytes=o["bytes"], tensor_pct=l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                   inst=l.get("smsp__inst_executed.sum"))
        out.append(rec)
        a = agg.setdefault(o["op_class"], dict(launches=0, ns=0.0, dram_bytes=0.0, algorithmic_bytes=0))
        a["launches"] += 1; a["ns"] += t; a["dram_bytes"] += dram; a["algorithmic_bytes"] += o["bytes"]
    total_ns = sum(r["ns"] for r in out if r["index"] >= 0)  # one reverse step (per-chain ops, index < 0, are listed but not part of it)
    for a in agg.values():
        a["share_of_step"] = a["ns"] / total_ns
        a["dram_bytes_per_launch"] = a["dram_bytes"] / a["launches"]
        a["us_per_launch"] = a["ns"] / a["launches"] / 1e3
    json.dump(dict(note="ncu per-launch times are cold-cache and serialised: compare SHARES with bench.py, not absolutes",
                   step_us=total_ns / 1e3, per_class=dict(sorted(agg.items(), key=lambda kv: -kv[1]["ns"])), launches=out), open(out_json, "w"), indent=1)
    print(f"{n} launches, step {total_ns / 1e3:.1f} us under ncu")
    for k, a in list(sorted(agg.items(), key=lambda kv: -kv[1]["ns"]))[:8]:
        print(f"  {k}: share {a['share_of_step']:.3f}, {a['us_per_launch']:.1f} us/launch, dram {a['dram_bytes_per_launch'] / 1e6:.1f} MB vs algorithmic "
              f"{a['algorithmic_bytes'] / a['launches'] / 1e6:.1f} MB")


def full(rep, out_json):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        out.append({k: (r[ix[k]] + (" " + units[ix[k]] if units[ix[k]] else "")) for k in keep if k in ix})
    json.dump(out, open(out_json, "w"), indent=1)
    for o in out:
        print({k: o[k] for k in ("Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") if k in o})


if __name__ == "__main__":
    {"step": step, "full": full}[sys.argv[1]](*sys.argv[2:])
