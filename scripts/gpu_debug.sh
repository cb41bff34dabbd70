cd $GRAFT_REPO_ROOT
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py --smoke 2>&1 | grep -v "^$" | head -40
