# scratch helper for gpurun calls (last use: final validation)
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
