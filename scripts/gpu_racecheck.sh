cd $GRAFT_REPO_ROOT
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 8 python __graft_entry__.py --smoke 2>&1 | grep -v "^$" | tail -25
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_naive.py -x -q 2>&1 | tail -4
