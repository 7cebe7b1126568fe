#!/usr/bin/env python
"""Condense an ncu report (read here, no GPU needed) into the few numbers the roofline talks about.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_full_summary.txt
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/smem % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (active)"),
    ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "bf16 tensor ops % of peak"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "TMEM cycles active %"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (active)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math pipe"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall mio throttle"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def traffic_json(rows, hdr, idx, units, batch):
    """{stage: dram bytes per image} from dram__bytes_read.sum + dram__bytes_write.sum of the FIRST launch of every
    kernel of the step (bench.py reports it as roofline.traffic)."""
    import json
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    stage_of = (("preprocess", "preprocess"), ("conv1_kernel", "conv1"), ("conv3x3_pair", "conv2"),
                ("conv3x3_kernel<32", "conv2"), ("conv3x3_kernel<64", "conv3"), ("conv3x3_stream", "conv4"),
                ("linear_splitk", "fc1"), ("head_tail", "tail"))
    out = {}
    for row in rows:
        name = row[idx["Kernel Name"]]
        for key, stage in stage_of:
            if key in name and stage not in out:
                total = 0.0
                for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(row[idx[m]].replace(",", "")) * scale.get(units[idx[m]], 1.0)
                out[stage] = total / batch
    return json.dumps({"batch": batch, "source": "ncu --set full, first launch of each kernel", "dram_bytes_per_image": out},
                      indent=1)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic":
        print(traffic_json(rows[2:], hdr, idx, units, int(sys.argv[3])))
        return
    for row in rows[2:]:
        name = row[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        print(f"== {name}  grid {row[idx['Grid Size']]} block {row[idx['Block Size']]}")
        for key, label in WANT:
            if key in idx:
                print(f"   {label:34s} {row[idx[key]]:>16s} {units[idx[key]]}")
        print()


if __name__ == "__main__":
    main()
