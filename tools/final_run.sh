# Round-end measurement on one GPU: full GPU test suite, default bench line, launch list and one ncu --set full step.
# usage (gpurun): bash tools/final_run.sh TAG
tag=${1:-final}
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_tests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python tools/bench_summary.py gpurun_out/${tag}_bench.json 2>/dev/null | head -30
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-graph --timed-only > gpurun_out/${tag}_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"x3_|knn_k32" --launch-skip 81 -c 27 -o gpurun_out/${tag}_x3 -f \
  python bench.py --steps 2 --warmup 3 --no-graph --timed-only > gpurun_out/${tag}_x3.log 2>&1; echo "ncu full rc=$?"
