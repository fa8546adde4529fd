#!/usr/bin/env python
"""Print the key metrics of an ncu --set full capture (run in the build
container: ncu -i reads .ncu-rep files without a GPU).
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [out.csv]"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'lts__t_sector_hit_rate.pct']


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    names = [r[hdr.index('Kernel Name')].split('(')[0] for r in rows[2:]]
    print("kernels:", names)
    keep = []
    for i, h in enumerate(hdr):
        if h in WANT or 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
            vals = [r[i] for r in rows[2:]]
            keep.append((h, units[i], vals))
    for h, u, vals in keep:
        if 'pcsamp' in h:
            continue
        print("%-70s %-14s %s" % (h, u, vals))
    st = sorted(((h, vals) for h, u, vals in keep if 'pcsamp' in h),
                key=lambda x: -float(x[1][0] or 0))
    tot = sum(float(v[0] or 0) for _, v in st)
    print("stall samples (first launch), total", tot)
    for h, vals in st[:10]:
        print("   %-55s %8s  %5.1f%%" % (h.replace('smsp__pcsamp_warps_issue_stalled_', ''),
                                          vals[0], 100 * float(vals[0] or 0) / tot))
    if len(sys.argv) > 2:
        with open(sys.argv[2], 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['metric', 'unit'] + names)
            for h, u, vals in keep:
                w.writerow([h, u] + vals)


if __name__ == "__main__":
    main()
