"""Print the metrics that matter from an .ncu-rep (raw page): used to write profiles/*.txt

    python tools/ncu_summary.py <file.ncu-rep>                      metrics of every captured launch
    python tools/ncu_summary.py <file.ncu-rep> --traffic KEY ALGORITHMIC_BYTES [SOURCE]
        additionally records dram__bytes_read.sum + dram__bytes_write.sum of the first launch under KEY in
        profiles/ncu_traffic.json, the tracked file bench.py takes `roofline.traffic` from (no literal in bench.py).
"""
import csv, json, os, subprocess, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed.sum','smsp__inst_executed.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_issued.avg.per_cycle_active',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'launch__occupancy_limit_warps','launch__waves_per_multiprocessor','launch__grid_size','sm__cycles_elapsed.avg','sm__cycles_active.avg',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
 'lts__t_bytes.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__cycles_active.avg',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_selected_per_issue_active.ratio',
 'smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active']
out = subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
H,U = rows[0],rows[1]
for V in rows[2:]:
    print('kernel:', V[H.index('Kernel Name')] if 'Kernel Name' in H else '?')
    for w in WANT:
        if w in H:
            i=H.index(w); print(f'  {w:92s} {V[i]:>16s} {U[i]}')


def _bytes(v, unit):
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(v.replace(",", "")) * mult


if "--traffic" in sys.argv:
    a = sys.argv.index("--traffic")
    key, algo = sys.argv[a + 1], float(sys.argv[a + 2])
    source = sys.argv[a + 3] if len(sys.argv) > a + 3 else os.path.basename(sys.argv[1])
    V = rows[2]
    rd = _bytes(V[H.index("dram__bytes_read.sum")], U[H.index("dram__bytes_read.sum")])
    wr = _bytes(V[H.index("dram__bytes_write.sum")], U[H.index("dram__bytes_write.sum")])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(path))
    except OSError:
        d = {}
    d[key] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr, "algorithmic_bytes": algo,
              "ratio": (rd + wr) / algo, "source": source}
    json.dump(d, open(path, "w"), indent=1)
    print("traffic", key, d[key])
