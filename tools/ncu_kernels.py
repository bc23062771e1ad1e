"""One table row per kernel launch of an `ncu --set full` report that holds many launches (development only).

    python tools/ncu_kernels.py rep.ncu-rep [> profiles/rN_ncu_kernels.txt]

Prints, per captured launch: duration, DRAM bytes read + written, achieved DRAM bandwidth against the measured copy
peak (MEASURED_PEAKS.json), issue-slot and L1 data-pipe utilisation, achieved occupancy, registers, the dominant
stall reason. The first launch of every distinct kernel is marked with '*'.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6545.9

    def val(r, k, default=float("nan")):
        try:
            return float(r[ix[k]].replace(",", ""))
        except (KeyError, ValueError):
            return default

    def to_bytes(r, k):
        v = val(r, k)
        u = units[ix[k]].lower() if k in ix else ""
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)

    def to_ms(r, k):
        v = val(r, k)
        u = units[ix[k]].lower() if k in ix else ""
        return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3, "nsecond": 1e-6}.get(u, 1)

    stall_keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print(f"# {os.path.basename(rep)}: ncu --set full --clock-control none; HBM peak {peak} GB/s (MEASURED_PEAKS.json)")
    print(f"{'kernel':44s} {'grid':>9s} {'blk':>4s} {'regs':>4s} {'ms':>8s} {'dramGB':>8s} {'GB/s':>7s} {'ofpeak':>6s} "
          f"{'issue%':>6s} {'L1pipe%':>7s} {'occ%':>5s} {'L2hit%':>6s}  top stall (warps per issue)")
    seen = set()
    for r in data:
        name = r[ix["Kernel Name"]]
        short = name.replace("fdn::", "").replace("(int)", "").replace("void ", "")
        short = short.split("(")[0] if "<" not in short else short[:short.index(">") + 1]
        ms = to_ms(r, "gpu__time_duration.sum")
        by = to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
        gbs = by / ms / 1e6 if ms > 0 else float("nan")
        stalls = sorted(((val(r, k, 0.0), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                         for k in stall_keys), reverse=True)[:2]
        mark = "*" if short not in seen else " "
        seen.add(short)
        print(f"{mark}{short[:43]:43s} {int(val(r, 'launch__grid_size', 0)):9d} {int(val(r, 'launch__block_size', 0)):4d} "
              f"{int(val(r, 'launch__registers_per_thread', 0)):4d} {ms:8.4f} {by / 1e9:8.4f} {gbs:7.0f} {gbs / peak:6.3f} "
              f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{val(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):7.1f} "
              f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} {val(r, 'lts__t_sector_hit_rate.pct'):6.1f}  "
              + ", ".join(f"{n} {v:.2f}" for v, n in stalls))


if __name__ == "__main__":
    main()
