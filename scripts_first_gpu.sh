set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
python -m pytest tests -x -q -m gpu 2>&1 | tail -30
