"""Summarise an ncu report (read with `ncu -i … --page raw --csv`) into a small JSON: one entry per captured launch with
the metrics the roofline discussion uses.  Usage: python scripts/ncu_extract.py <report.ncu-rep> <out.json>"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_sha():
    """Hash of the kernel sources: bench.py only quotes a capture's DRAM traffic while the kernels are unchanged."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "meg-decoding_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read_bytes",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "tensor_smem_wavefronts_pct",
    # the tensor-pipe utilisation north_star asks for (ncu's own figure, next to the fraction of the measured cuBLAS peak)
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_hmma_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_pipe_instructions",
    "smsp__cycles_active.avg": "smsp_cycles_active",
    "l1tex__m_l1tex2xbar_write_bytes.sum": "sm_to_l2_write_bytes",
}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3,
        "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        e = {"kernel": r[hdr.index("Kernel Name")][:120]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                e[KEYS[h]] = v * UNIT.get(units[i], 1.0)
        launches.append(e)
    json.dump({"report": rep.split("/")[-1], "source_sha": source_sha(), "launches": launches}, open(out, "w"), indent=1)
    for e in launches:
        print(e["kernel"][:60], {k: (round(v, 1) if isinstance(v, float) else v) for k, v in e.items() if k != "kernel"})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
