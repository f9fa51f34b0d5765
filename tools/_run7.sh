python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_v6.log 2>&1
python bench.py --steps 2 --warmup 3 --no-also > gpurun_out/r02_bench_noalso.json 2> gpurun_out/r02_bench_noalso.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02f_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-also > gpurun_out/r02f_ncu_l.log 2>&1
python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc --launch-skip 60 --launch-count 4 -o gpurun_out/r02f_prof_tc python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/r02f_ncu_f.log 2>&1
NRT_PROF_REPS=2 python tools/prof_kernels.py sdf_infer > /dev/null 2>&1 && \
NRT_PROF_REPS=2 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc --launch-skip 4 --launch-count 4 -o gpurun_out/r02f_prof_sdf python tools/prof_kernels.py sdf_infer > gpurun_out/r02f_ncu_s.log 2>&1
ls -la gpurun_out/*.ncu-rep
exit 0
