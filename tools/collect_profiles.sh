#!/bin/bash
# Copies the judged artefacts of a tools/gpu_round.sh / gpu_r2b.sh run from gpurun_out/ into profiles/ under a tag.
# Usage (here, after the gpurun call): tools/collect_profiles.sh <tag>
t=$1
cp gpurun_out/bench_$t.json profiles/${t}_bench.json
cp gpurun_out/conv_events_$t.txt profiles/${t}_conv_layers_cuda_events.txt
cp gpurun_out/conv_events_$t.txt.plan profiles/${t}_plan_B32_512.txt
cp gpurun_out/conv_events_tuning_$t.txt.tune profiles/${t}_autotune_candidates.txt
cp gpurun_out/conv_layers_$t.txt profiles/${t}_conv_layers_ncu.txt
[ -f gpurun_out/parity_report.txt ] && cp gpurun_out/parity_report.txt profiles/${t}_parity_report.txt
python tools/launch_table.py gpurun_out/launches_$t.csv > profiles/${t}_launches_bench_steps2_warmup1.txt
python tools/launch_summary.py gpurun_out/launches_$t.csv $t 2 > /dev/null
for s in 0 5 58 63; do
  [ -f gpurun_out/src_roles_${t}_$s.txt ] && cp gpurun_out/src_roles_${t}_$s.txt profiles/${t}_ncu_roles_conv$s.txt
  [ -f gpurun_out/src_${t}_$s.source.csv ] && python tools/src_top.py gpurun_out/src_${t}_$s.source.csv 30 1 > profiles/${t}_ncu_source_top_conv$s.txt
done
python - "$t" <<'PY'
import csv, json, sys
t = sys.argv[1]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'sm__cycles_elapsed.avg.per_second',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_dim_x',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'l1tex__m_xbar2l1tex_read_bytes.sum']
plan = [l for l in open('gpurun_out/conv_events_%s.txt.plan' % t) if l.startswith('gemm ')]
for s in (0, 5, 58, 63):
    try:
        rows = list(csv.reader(open('gpurun_out/src_%s_%d.raw.csv' % (t, s))))
    except OSError:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: {'value': v, 'unit': u} for h, u, v in zip(hdr, units, vals) if h in keep}
    json.dump({'capture': 'ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_gemm -s %d -c 1 '
                          'python tools/profile_forward.py --batch 32 --size 512 --iters 1' % s,
               'layer': plan[s].strip() if s < len(plan) else '', 'metrics': d},
              open('profiles/%s_ncu_full_conv%d.json' % (t, s), 'w'), indent=1)
PY
ls profiles | grep $t
