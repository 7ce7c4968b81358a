set -x
cd $GRAFT_REPO_ROOT
python __graft_entry__.py smoke 2>&1 | tail -30
