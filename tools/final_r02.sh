# round-2 final single-GPU batch: tests, smoke, both bench arms, launch list, full ncu capture of the column kernel
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_gputest_final.log; cat gpurun_out/r02_gputest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
python bench.py --impl reference > gpurun_out/bench_r02f_ref.json 2> gpurun_out/bench_r02f_ref.err
python bench.py > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
cut -c1-300 gpurun_out/bench_r02f.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:letkf_nsp -s 1 -c 1 -f -o gpurun_out/r02_nsp_f python tools/prof_case.py 128 80 60 > gpurun_out/ncu_f.log 2>&1
