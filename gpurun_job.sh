cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
