#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: samples per SASS region + hottest instructions,
and the headline raw metrics.  usage: tools/ncu_roles.py SOURCE.csv [RAW.csv]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S = lambda r: int(r[ix['# Samples']] or 0)
tot = sum(S(r) for r in data)
print("total samples", tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
top = sorted(range(len(data)), key=lambda i: -S(data[i]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(i, S(r), r[ix['Instructions Executed']], r[ix['Source']].strip()[:80], st)
if len(sys.argv) > 2 and sys.argv[2] != '-':
    rr = list(csv.reader(open(sys.argv[2])))
    h, u, r = rr[0], rr[1], rr[2]
    for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
              "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg", "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg",
              "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
              "launch__registers_per_thread", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"]:
        if k in h:
            print(k, r[h.index(k)], u[h.index(k)])
