#!/usr/bin/env python
"""profiles/r02_traffic.json from an ncu --set full capture of one build (regenerate after every
kernel change; bench.py quotes it as roofline.traffic with the source named).

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep "<what was profiled>" > profiles/r02_traffic.json
"""
import csv
import io
import json
import re
import subprocess
import sys

PHASE_OF = {"k_tbin_": "binning", "k_plan_": "tile_plan", "k_tile_count": "count_pass",
            "k_tile_fill": "fill_pass", "k_sorted_dst": "offset_scan", "k_max_and_sum": "offset_scan"}


def main():
    rep, what = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]

    def val(r, name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)

    per_kernel, per_phase, other = {}, {}, {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        m = re.search(r"\b(k_\w+)", name)
        short = m.group(1) if m else name.split("(")[0]
        b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        per_kernel.setdefault(short, []).append(b)
        for key, ph in PHASE_OF.items():
            if key in short:
                per_phase.setdefault(ph, []).append((short, b))
        if "k_tile_count" in short:
            g = lambda m: float(r[hdr.index(m)]) if m in hdr else None
            other = {
                "kernel": "k_tile_count",
                "l1tex_throughput_pct": g("l1tex__throughput.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "tensor_pipe_active_pct": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "fp64_pipe_active_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                "warp_instructions": g("smsp__inst_executed.sum"),
            }
    # one build = one launch of every kernel: average repeated launches of the same kernel
    avg = {k: sum(v) / len(v) for k, v in per_kernel.items()}
    phases = {}
    for ph, items in per_phase.items():
        names = sorted(set(n for n, _ in items))
        phases[ph] = sum(avg[n] for n in names)
    phases["step"] = sum(avg.values())
    json.dump({"source": f"{rep}: ncu --set full --clock-control none ({what}); dram__bytes_read.sum + "
                         "dram__bytes_write.sum per launch",
               "dram_bytes_per_launch": phases, "per_kernel": avg, "other_units": other},
              sys.stdout, indent=1)
    sys.stdout.write("\n")


if __name__ == "__main__":
    main()
