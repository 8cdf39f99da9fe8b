"""Summarises an .ncu-rep (raw page + source page) into a compact JSON/text (run where ncu is installed)."""
import csv, json, subprocess, sys, collections, io

rep = sys.argv[1]
out_json = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hdr, units, vals = rows[0], rows[1], rows[2 + which]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'smsp__sass_thread_inst_executed_op_dmma_pred_on.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'smsp__inst_executed.sum']
keys += [k for k in d if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio')]
keys += [k for k in d if 'pipe_tensor' in k and k not in keys]
summ = {k: {'unit': d[k][0], 'value': d[k][1]} for k in keys if k in d}
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']]) for r in data)
samp = collections.Counter(); ex = collections.Counter()
for r in data:
  s = r[ix['Source']].strip()
  op = s.split()[1] if s.startswith('@') else s.split()[0]
  samp[op] += int(r[ix['# Samples']]); ex[op] += int(r[ix['Instructions Executed']])
summ['_stall_samples_by_opcode'] = {op: {'pct': round(100 * c / tot, 2), 'executed': ex[op]} for op, c in samp.most_common(12)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
top = sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:12]
summ['_top_instructions'] = [{'samples': int(r[ix['# Samples']]), 'sass': r[ix['Source']].strip()[:80],
                              'stalls': dict(sorted({c: int(r[ix[c]]) for c in stall_cols if int(r[ix[c]]) > 0}.items(), key=lambda kv: -kv[1])[:3])} for r in top]
print(json.dumps(summ, indent=1))
if out_json:
  json.dump(summ, open(out_json, 'w'), indent=1)
