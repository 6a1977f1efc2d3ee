import csv, sys
from collections import defaultdict
for n in sys.argv[1:]:
    rows=[r for r in csv.reader(open(f'gpurun_out/ncu_l2_{n}.csv')) if len(r)>10]
    hdr=rows[0]; ci={h:i for i,h in enumerate(hdr)}
    d=defaultdict(dict)
    for r in rows[1:]:
        d[(r[ci['ID']], r[ci['Kernel Name']][:60], r[ci['Grid Size']])][r[ci['Metric Name']]]=r[ci['Metric Value']]
    print(n)
    for k,v in d.items():
        f=lambda x: float(v[x].replace(',',''))
        print(k[1][-26:],k[2], f"t={f('gpu__time_duration.sum')/1e3:.0f}us dramR={f('dram__bytes_read.sum')/1e6:.0f}MB l2hit={f('lts__t_sector_op_read_hit_rate.pct'):.0f}% texRdHit={f('lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum')/1e6:.1f}M miss={f('lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum')/1e6:.1f}M allmiss={f('lts__t_sectors_srcunit_tex_lookup_miss.sum')/1e6:.1f}M l1miss={f('l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum')/1e6:.1f}M issue={f('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f}% inst={f('smsp__inst_executed.sum')/1e6:.0f}M")
