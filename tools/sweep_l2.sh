for wl in c2 rgb; do for m in 3 0 12 8 4 9 6; do
echo -n "$wl hints=$m : "; VN_BULK_L2_HINTS=$m python bench.py --workload $wl --steps 4000 --warmup 50 --quick 2>&1 | tail -1 | cut -c1-120
done; done
