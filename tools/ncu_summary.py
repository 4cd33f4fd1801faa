"""Summarise an `ncu --set full` report into the small JSON kept under profiles/ (runs here, no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-name-substring] > profiles/rNN_<kernel>.json

One entry per profiled launch whose name contains the substring: duration, DRAM bytes (the `roofline.traffic`
of the bench line), L1/L2 hit rates, the busiest units, occupancy, registers, shared-memory wavefronts."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "l1tex__t_sector_hit_rate.pct": "l1tex_hit_rate_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_to_l1_read_sectors",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "l1_global_load_sectors",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts.avg": "data_pipe_wavefronts_per_sm",
    "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg": "data_pipe_wavefronts_global_per_sm",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg": "data_pipe_wavefronts_shared_per_sm",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_l1_fill_bytes",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "sm__inst_executed.sum": "warp_instructions",
    "sm__cycles_elapsed.max": "elapsed_cycles",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__ops_path_tensor_src_tf32_dst_fp32.avg.per_cycle_elapsed": "tf32_flop_per_clk_per_sm",
    "sm__ops_path_tensor_src_tf32_dst_fp32.avg.pct_of_peak_sustained_elapsed": "tf32_pct_of_ncu_peak",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_memory_path_active_pct",
}


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {}
    for i, name in enumerate(head):
        for metric, key in WANT.items():
            if name == metric or name.endswith("." + metric):
                col.setdefault(key, (i, units[i]))
    name_i = head.index("Kernel Name")
    out = []
    for r in body:
        if pat and pat not in r[name_i]:
            continue
        e = {"kernel": r[name_i], "grid": r[head.index("Grid Size")], "block": r[head.index("Block Size")]}
        for key, (i, unit) in col.items():
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            e[key + ("_" + unit.replace("/", "_per_") if unit and unit not in ("%", "sector", "inst", "cycle", "register/thread") else "")] = v
        if "dram_bytes_read_Mbyte" in e or "dram_bytes_read_byte" in e or "dram_bytes_read_Gbyte" in e:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = 0.0
            for k in list(e):
                if k.startswith("dram_bytes_"):
                    tot += e[k] * scale.get(k.rsplit("_", 1)[1], 1.0)
            e["dram_bytes_per_launch"] = tot
        out.append(e)
    res = {"source": rep, "launches": out}
    # `--hybrid k_tc_pack,k_tc_mma,k_spmm`: DRAM bytes of ONE hybrid propagation = the first launch of each named kernel
    # after the first k_tc_pack (bench.py reads `dram_bytes_per_launch` as roofline.traffic)
    if "--hybrid" in sys.argv:
        names = sys.argv[sys.argv.index("--hybrid") + 1].split(",")
        start = next((i for i, e in enumerate(out) if names[0] in e["kernel"]), None)
        if start is not None:
            tot, used, dur = 0.0, [], 0.0
            j = start
            for nm in names:
                while j < len(out) and nm not in out[j]["kernel"]:
                    j += 1
                if j < len(out):
                    tot += out[j].get("dram_bytes_per_launch", 0.0)
                    dur += next((v for k, v in out[j].items() if k.startswith("duration")), 0.0)
                    used.append(out[j]["kernel"][:60])
                    j += 1
            res["dram_bytes_per_launch"] = tot
            res["hybrid_kernels"] = used
            res["hybrid_duration_sum_us_under_ncu"] = dur
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
