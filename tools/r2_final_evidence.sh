#!/bin/bash
# Round-2 final evidence on one B200: GPU tests, smoke, default bench, ncu launch list of the bench command, ncu --set full
# of one timestep's conv launches, every-kernel table, config-4 latency.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3) > gpurun_out/fin_tests.log; cat gpurun_out/fin_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.txt 2>&1; tail -2 gpurun_out/fin_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/fin_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roofline'].get('frac_of_tf32_pipe'),d['roofline']['frac_of_step'],d['parity']['max_rel_err'],d['cpu_baseline']['value']); t=d['train']; print({k:t[k] for k in t if k in ('maps_per_s','ms_per_step','loss')}, t['wgrad']['achieved_tflops'], t['conv_fwd_dgrad']['achieved_tflops'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin_bench_ref.json 2> gpurun_out/fin_bench_ref.err; cut -c1-300 gpurun_out/fin_bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-train > gpurun_out/fin_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/fin_launches.csv > gpurun_out/fin_launches_summary.txt 2>&1; head -25 gpurun_out/fin_launches_summary.txt
RAMNET_FRONT_STREAM=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo -s 68 -c 34 -o gpurun_out/fin_halo_timestep -f python bench.py --no-graphs --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-train > gpurun_out/fin_ncu_halo.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/fin_halo_timestep.ncu-rep > gpurun_out/fin_halo_timestep.txt 2>&1; grep -c "^==" gpurun_out/fin_halo_timestep.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/fin_ncu_all.csv python tools/ncu_all_kernels.py > gpurun_out/fin_ncu_all.log 2>&1; echo "ncu all rc=$?"
python tools/ncu_kernel_table.py gpurun_out/fin_ncu_all.csv --skip-first-half > gpurun_out/fin_ncu_all.txt 2>&1; head -12 gpurun_out/fin_ncu_all.txt
timeout 300 python tools/latency_bench.py --json gpurun_out/fin_latency.json > gpurun_out/fin_latency.txt 2>&1; tail -4 gpurun_out/fin_latency.txt
rm -f gpurun_out/fin_halo_timestep.ncu-rep.tmp; ls -la gpurun_out/fin_halo_timestep.ncu-rep
