"""Curated one-screen summary of an `ncu --set full` report (the numbers DESIGN.md / bench.py quote):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...]      (needs the ncu CLI; reads, does not profile)"""
import csv, io, subprocess, sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('sm__cycles_elapsed.avg.per_second', 'SM clock'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput (% of peak)'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput (% of peak)'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput (% of peak)'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active (% of active cycles)'),
    ('sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'tcgen05 bf16 MMA ops (% of peak, elapsed)'),
    ('sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'mma.sync bf16 ops (% of peak, elapsed)'),
    ('sm__inst_executed_pipe_uniform.sum', 'uniform-pipe instructions'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy'),
    ('launch__registers_per_thread', 'registers / thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem / block'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
]
STALLS = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio'


def main():
    for path in sys.argv[1:]:
        txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            print('## %s  (%s)' % (r[ix['Kernel Name']][:110], path.split('/')[-1]))
            for k, label in KEYS:
                if k in ix and r[ix[k]] != '':
                    print('  %-46s %14s %s' % (label, r[ix[k]], units[ix[k]]))
            st = []
            for h, i in ix.items():
                if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
                    try:
                        st.append((float(r[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            print('  top stalls (warps per issue-active cycle):  ' + ', '.join('%s %.2f' % (n, v) for v, n in st[:6]))
            print()


if __name__ == '__main__':
    main()
