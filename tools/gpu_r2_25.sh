set -x
O=gpurun_out/r2; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > $O/t25_all.log 2>&1; echo "rc=$?" >> $O/t25_all.log
timeout 900 python bench.py > $O/bench25.json 2> $O/bench25.err; echo "rc=$?" >> $O/bench25.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke25.log 2>&1; echo "rc=$?" >> $O/smoke25.log
