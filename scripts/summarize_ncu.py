"""Turn the ncu exports of one profiled recognize() step into the committed per-layer evidence.

    python scripts/summarize_ncu.py <launches_metrics.csv> <conv_raw.csv> <out_prefix>

* ``launches_metrics.csv``: ``ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
  lts__t_bytes.sum,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active,... --csv`` (long format,
  every launch of the step).
* ``conv_raw.csv``: ``ncu -i conv.ncu-rep --page raw --csv`` of an ``--set full -k regex:conv_tc`` capture (wide format).
Writes ``<out_prefix>_layers.csv`` (one row per launch: kernel, layer, ms, DRAM read/write GB, L2 GB, tensor-pipe %, TFLOP/s
from the algorithmic FLOPs) and ``<out_prefix>_traffic.json`` (DRAM bytes per conv_tc launch, what bench.py reports as
``roofline.traffic``)."""
import csv
import json
import re
import sys

# launch order of conv_tc_kernel in one recognize() step (api.cu: b2o_craft_forward, b2o_crnn_forward), with the
# algorithmic FLOPs per unit: (name, pixels-per-image divisor relative to the detector input, taps*cin, cout)
CRAFT = [("stem", 1, 27, 64), ("slice1.3", 1, 576, 64), ("slice1.7", 4, 576, 128), ("slice1.10", 4, 1152, 128),
         ("slice2.14", 16, 1152, 256), ("slice2.17", 16, 2304, 256), ("slice3.20", 16, 2304, 256), ("slice3.24", 64, 2304, 512),
         ("slice3.27", 64, 4608, 512), ("slice4.30", 64, 4608, 512), ("slice4.34", 256, 4608, 512), ("slice4.37", 256, 4608, 512),
         ("slice5.1", 256, 4608, 1024), ("slice5.2", 256, 1024, 1024), ("upconv1.0", 256, 1536, 512), ("upconv1.3", 256, 4608, 256),
         ("upconv2.0", 64, 768, 256), ("upconv2.3", 64, 2304, 128), ("upconv3.0", 16, 384, 128), ("upconv3.3", 16, 1152, 64),
         ("upconv4.0", 4, 192, 64), ("upconv4.3", 4, 576, 32), ("conv_cls.0", 4, 288, 32), ("conv_cls.2", 4, 288, 32),
         ("conv_cls.4", 4, 288, 16)]
# with the commuted decoder upsampling (default since r2e) upconv2/3/4.conv.0 run as a low-resolution ".y" GEMM + a ".s" layer
CRAFT_COMMUTED = []
for _name, _div, _k, _co in CRAFT:
    _split = {"upconv2.0": 256, "upconv3.0": 128, "upconv4.0": 64}.get(_name)
    if _split:
        CRAFT_COMMUTED += [(_name + ".y", _div * 4, _split, _co), (_name + ".s", _div, _k - _split, _co)]
    else:
        CRAFT_COMMUTED.append((_name, _div, _k, _co))
CRNN = [("conv_2", 6200, 576, 128), ("conv_3", 6200, 1152, 256), ("conv_4", 1500, 2304, 256), ("conv_5", 1500, 2304, 512),
        ("conv_6", 350, 4608, 512), ("conv_7", 350, 4608, 512), ("stn.conv_a", 350, 12800, 16), ("stn.conv_b", 350, 400, 32),
        ("stn.dense_a", 1, 11200, 64), ("fc_9", 50, 3584, 128), ("lstm_in_1", 50, 128, 1024), ("lstm_in_2", 50, 128, 1024)]


def short(name):
    m = re.search(r"(\w+)(<[^(]*>)?\(", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def main():
    metrics_csv, raw_csv, prefix = sys.argv[1:4]
    pages = int(sys.argv[4]) if len(sys.argv) > 4 else 32
    px = 1536 * 1536
    launches = {}
    with open(metrics_csv) as f:
        rows = [r for r in csv.reader(f) if r and r[0].isdigit()]
    for r in rows:
        d = launches.setdefault(int(r[0]), {"kernel": short(r[4]), "grid": r[8]})
        try:
            d[r[12]] = float(r[14].replace(",", ""))
        except ValueError:                         # "n/a": the metric does not exist on this chip
            pass
    order = [launches[k] for k in sorted(launches)]
    conv = [d for d in order if d["kernel"].startswith("conv_tc_kernel")]
    crops = None
    craft = CRAFT_COMMUTED if len(conv) == len(CRAFT_COMMUTED) + len(CRNN) else CRAFT
    for i, d in enumerate(conv):
        if i < len(craft):
            n, div, k, co = craft[i]
            d["layer"], d["flop"] = "craft." + n, 2.0 * pages * px / div * k * co
        else:
            n, per, k, co = CRNN[i - len(craft)]
            if crops is None:                      # crops = grid-independent: recover from the DRAM-free fact that conv_2 is per crop
                crops = int(sys.argv[5]) if len(sys.argv) > 5 else 1028
            d["layer"], d["flop"] = "crnn." + n, 2.0 * crops * per * k * co
    out = []
    for d in order:
        ms = d.get("gpu__time_duration.sum", 0.0) / 1e6
        rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        out.append({"kernel": d["kernel"], "layer": d.get("layer", ""), "grid": d["grid"], "ms": round(ms, 4),
                    "dram_read_GB": round(rd / 1e9, 4), "dram_write_GB": round(wr / 1e9, 4),
                    "dram_TBps": round((rd + wr) / ms / 1e9, 3) if ms else 0,
                    "l2_GB": round(d.get("lts__t_bytes.sum", 0.0) / 1e9, 3),
                    "tensor_pipe_pct": round(d.get("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", 0.0), 2),
                    "tensor_inst": int(d.get("sm__inst_executed_pipe_tensor.sum", 0)),
                    "alg_TFLOP": round(d.get("flop", 0.0) / 1e12, 4),
                    "TFLOPps": round(d.get("flop", 0.0) / ms / 1e9, 1) if ms and "flop" in d else ""})
    with open(prefix + "_layers.csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(out[0]))
        w.writeheader()
        w.writerows(out)
    conv_rows = [o for o in out if o["kernel"].startswith("conv_tc_kernel")]
    total_ms = sum(o["ms"] for o in out)
    summary = {
        "source": metrics_csv, "launches": len(out), "kernel_ms_total": round(total_ms, 3),
        "conv_tc_launches": len(conv_rows), "conv_tc_ms": round(sum(o["ms"] for o in conv_rows), 3),
        "conv_tc_share_of_kernel_time": round(sum(o["ms"] for o in conv_rows) / total_ms, 4),
        "conv_tc_dram_bytes_per_step": sum(o["dram_read_GB"] + o["dram_write_GB"] for o in conv_rows) * 1e9,
        "conv_tc_dram_bytes_per_launch": sum(o["dram_read_GB"] + o["dram_write_GB"] for o in conv_rows) * 1e9 / max(len(conv_rows), 1),
        "conv_tc_l2_bytes_per_step": sum(o["l2_GB"] for o in conv_rows) * 1e9,
        "conv_tc_alg_TFLOP_per_step": round(sum(o["alg_TFLOP"] for o in conv_rows), 3),
        "note": "ncu per-launch times are cold-cache and serialised at unthrottled clocks: compare shares, not absolutes",
    }
    # the --set full capture: a few columns per conv launch
    try:
        with open(raw_csv) as f:
            raw = list(csv.reader(f))
        hi = next(i for i, r in enumerate(raw) if r and r[0] == "ID")
        hdr, data = raw[hi], raw[hi + 2:]
        keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
                "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
                "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
        keep = [k for k in keep if k in hdr]
        with open(prefix + "_conv_full.csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["launch", "layer", "kernel"] + keep)
            for i, r in enumerate(data):
                layer = conv[i]["layer"] if i < len(conv) else ""
                w.writerow([i, layer, short(r[hdr.index("Kernel Name")])] + [r[hdr.index(k)] for k in keep])
        summary["full_capture"] = raw_csv
    except (OSError, StopIteration) as exc:
        summary["full_capture"] = f"unavailable: {exc}"
    with open(prefix + "_traffic.json", "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
