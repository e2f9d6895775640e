for m in 0 1 2 4 3 5 6; do PB_DEBUG_MODE=$m python tests/analysis/kbench.py T:16 cfg5:16 --tag "debug mode $m"; done
