timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 100 python tests/analysis/kbench.py cfg4:1 T:1 cfg1:1 --tag "sep1 fixed-point blend"
