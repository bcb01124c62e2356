#!/usr/bin/env python3
"""profiles/<round>_ncu_summary.json from an `ncu --set full` capture of ONE k_chain launch (tools/round2_evidence.sh):
executed FP64 work, warp instructions, L2 and DRAM bytes of the launch, and the same per Markov-chain step.
bench.py reads the per-step figures for `roofline` (executed fraction, issue-slot fraction, L2 view).

  python tools/ncu_chain_summary.py gpurun_out/prof_r02_kchain.ncu-rep <steps in the captured launch> profiles/r02_ncu_summary.json [section]

`section` (default k_chain) names the entry; an existing output file is updated, so the single-chain capture (k_chain) and the
capture of the whole fleet in one launch (k_chain_fleet: one chain per SM, device-wide issue / FP64-pipe / L2 percentages) share a file."""
import os
import csv
import json
import subprocess
import sys


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [{h: (u, v) for h, u, v in zip(hdr, units, row)} for row in rows[2:]]


def num(cell):
    u, v = cell
    v = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "s": 1.0, "ns": 1e-9,
             "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "nsecond": 1e-9}
    return v * scale.get(u, 1.0)


def main():
    rep, steps, dst = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    section = sys.argv[4] if len(sys.argv) > 4 else "k_chain"
    launches = [r for r in raw_page(rep) if "k_chain" in r["Kernel Name"][1]]
    r = launches[-1]
    cyc = num(r["sm__cycles_elapsed.max"])
    per_cycle = lambda k: num(r[k]) * cyc   # noqa: E731
    dfma = per_cycle("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed")
    dadd = per_cycle("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed")
    dmul = per_cycle("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed")
    flop = 2 * dfma + dadd + dmul
    winst = num(r["smsp__inst_executed.sum"])
    dram = num(r["dram__bytes_read.sum"]) + num(r["dram__bytes_write.sum"])
    l2 = 32.0 * num(r["lts__t_sectors.sum"]) if "lts__t_sectors.sum" in r else None   # 32-byte sectors
    dur = num(r["gpu__time_duration.sum"])
    grid = r["Grid Size"][1] if "Grid Size" in r else None
    block = r["Block Size"][1] if "Block Size" in r else None
    def pct(k):
        return num(r[k]) if k in r else None
    out = {}
    if os.path.exists(dst):
        with open(dst) as f:
            out = json.load(f)
    out[section] = {
    "_": {
        "source": rep, "steps_in_launch": steps, "grid": grid, "block": block,
        "registers_per_thread": num(r["launch__registers_per_thread"]),
        "gpu_time_s": dur, "us_per_step_under_ncu": dur * 1e6 / steps,
        "thread_inst_dfma": dfma, "thread_inst_dadd": dadd, "thread_inst_dmul": dmul,
        "executed_fp64_flop": flop, "executed_fp64_flop_per_step": flop / steps,
        "warp_instructions": winst, "warp_instructions_per_step": winst / steps,
        "dram_bytes": dram, "dram_bytes_per_step": dram / steps,
        "l2_bytes": l2, "l2_bytes_per_step": (l2 / steps) if l2 else None,
        "issue_active_pct": num(r["smsp__issue_active.avg.pct_of_peak_sustained_active"]) if "smsp__issue_active.avg.pct_of_peak_sustained_active" in r else None,
        "device_wide_pct": {"sm_issue_active_of_elapsed": pct("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
                            "fp64_pipe_inst_of_active": pct("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                            "fp64_pipe_cycles_of_elapsed": pct("TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
                            "l2_throughput_of_elapsed": pct("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                            "l1tex_throughput_of_active": pct("l1tex__throughput.avg.pct_of_peak_sustained_active"),
                            "dram_throughput_of_elapsed": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")},
        "note": "workload S, `steps_in_launch` Markov-chain steps summed over the chains of the captured launch (one chain per CTA); "
                "counters are launch totals (per_cycle_elapsed x sm__cycles_elapsed.max for the FP64 thread-instruction counters); "
                "device_wide_pct are averages over all SMs / L2 slices of the device for that launch"}}
    out[section] = out[section]["_"]
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out[section]))


if __name__ == "__main__":
    main()
