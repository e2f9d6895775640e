K="timeout 120 python tests/analysis/kbench.py T:1 cfg1:1"
$K --tag "5 CTAs 48 regs"
cp photonbend_b200/libpbremap.so /tmp/lib_orig.so
cp scratch/lib_cta4.so photonbend_b200/libpbremap.so
$K --tag "4 CTAs 60 regs"
PB_SEP1_WAVES=2 $K --tag "4 CTAs waves2"
cp /tmp/lib_orig.so photonbend_b200/libpbremap.so
