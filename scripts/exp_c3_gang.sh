# pass B without the cache on C3-shaped input (the per-GPU share of bench.py's C3): CTA size of the direct variant
C3="--reads 12500000 --genome 12500000 --ksize 21 --sub-ppm 10000 --n-ppm 1000 --steps 2 --hint 280000000"
for th in 1024 768; do
  echo "== OXLI_B200_AGG_THREADS_DIRECT=$th"
  OXLI_B200_AGG_THREADS_DIRECT=$th timeout 100 python scripts/exp_part.py $C3 --digest --configs fused,auto
done
echo "== unhinted, fresh table every step"
timeout 100 python scripts/exp_part.py --reads 12500000 --genome 12500000 --ksize 21 --sub-ppm 10000 --n-ppm 1000 --steps 2 --hint 0 --fresh --digest --configs fused,auto
