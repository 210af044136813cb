#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` report of one FieldConv layer fwd+bwd at the cfg-2 layer shape
(tools/gpu_evidence.sh): per kernel of the layer (in launch order: aggregate, gemm_h_nn forward, aggregate_T, gemm_h_nn grouped,
pack_xhat_tn, gemm_h_tn) the DRAM bytes per launch and the pipe counters bench.py copies into `roofline`.
usage: make_ncu_traffic.py X.ncu-rep out.json source-note"""
import csv
import io
import json
import subprocess
import sys

KEYS = {"dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "us": "gpu__time_duration.sum"}
SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main():
    rep, dst, note = sys.argv[1], sys.argv[2], sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    order = ["aggregate", "gemm_h_nn", "aggregate_T", "gemm_h_nn_grouped", "pack_xhat_tn", "gemm_h_tn"]
    layer, total = {}, 0.0
    for name, r in zip(order, rows[2:]):
        ent = {"kernel": r[kn].split("(")[0].replace("void ", "")}
        for k, col in KEYS.items():
            i = hdr.index(col)
            ent[k] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
        ent["dram_bytes"] = ent.pop("dram_rd") + ent.pop("dram_wr")
        total += ent["dram_bytes"]
        layer[name] = {k: (round(v, 1) if isinstance(v, float) else v) for k, v in ent.items()}
    json.dump({"source": note, "cfg2_layer": layer, "layer_dram_bytes_fwd_bwd": total}, open(dst, "w"), indent=1)
    print(json.dumps(layer, indent=1), total)


if __name__ == "__main__":
    main()
