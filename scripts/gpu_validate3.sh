# cautious validation after the lost box: PatchMatch-phase tests one by one with hard timeouts, then the rest of the suite, progress in files
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for t in test_pm_phase_replay_small test_pm_phase_replay_right_view test_pm_phase_replay_r10 test_pm_phase_full_sweep_on_the_cones_crop test_pm_phase_cell_shard_two_ranks_in_one_process test_native_sweep_object_equals_the_python_schedule; do
  echo "== $t"; timeout 150 python -m pytest tests/test_gpu_pm.py -q -m gpu -k $t -s 2>&1 | grep -E "pm replay|passed|failed|error|Error|timeout" | head -5
done
echo "== smoke"; timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
echo "== rest of the suite"; timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_pm.py 2>&1 | tail -4
echo "== pm again x2"; for i in 1 2; do timeout 400 python -m pytest tests/test_gpu_pm.py -q -m gpu 2>&1 | tail -1; done
