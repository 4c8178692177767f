python -m pytest tests -m gpu -x -q > gpurun_out/t14.log 2>&1; tail -3 gpurun_out/t14.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 200 --warmup 5 > gpurun_out/b14.json 2> gpurun_out/b14.err; echo "bench rc=$?"
