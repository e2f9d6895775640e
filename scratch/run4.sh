M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_op_read.sum"
for band in 0 16; do
for wl in cfg5:16 T:16; do
echo "== band $band $wl"
PB_L2_AHEAD=0 PB_RASTER_BAND=$band ncu --metrics $M --clock-control none -k regex:remap_tiled -s 3 -c 1 python tests/analysis/kbench.py $wl --steps 2 --warmup 1 2>&1 | grep -E "dram__|lts__|gpu__time"
done; done
