set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -30
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
