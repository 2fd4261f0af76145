mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=line 2>&1 | grep -v "^$" | tail -25
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
