# A/B of the S2t kernels (register loads vs streamed ring; panel width / row-accumulator slots / ring size of the streamed one)
T="timeout -s KILL 120"
SCQED_DENSE_TILED=2 $T python tools/s2_rate.py 400:1184 500:1184 600:592 800:296 1000:296 1500:148 2000:148 3375:148 2>&1 | grep -v Warning
