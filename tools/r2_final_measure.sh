# Final single-GPU measurements of round 2 (run on the GPU box through gpurun): tests, the driver's bench command (its
# reference arm: profiles/r02_bench_reference_n1.json, unchanged since), the ncu launch list, and an ncu --set full capture
# of the traceback kernel.
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/pytest_final.log 2>&1; tail -3 gpurun_out/r2/pytest_final.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench_n1_final.json 2> gpurun_out/r2/bench_n1_final.err; tail -2 gpurun_out/r2/bench_n1_final.err
bash tools/r2_final_measure2.sh
ls gpurun_out/r2 | wc -l
