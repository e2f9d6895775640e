python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tests/analysis/kbench.py T:16 cfg5:16 T:1 cfg4:1 cfg1:1 cfg2:1 cfg3:1 --tag "$1"
