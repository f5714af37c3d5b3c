# eigh_batched matrices/s by size: streamed kernel with one / two CTAs per SM, register kernel
T="timeout -s KILL 120"
SCQED_DENSE_TILED=2 $T python tools/s2_rate.py 300:2072 400:1184 500:1184 600:592 700:592 800:592 2>&1 | grep -v Warning
SCQED_DENSE_TILED=2 SCQED_DTC_TWO=0 $T python tools/s2_rate.py 600:592 700:592 800:592 2>&1 | grep -v Warning
