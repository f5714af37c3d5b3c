# eigh_batched matrices/s by size (auto path)
T="timeout -s KILL 120"
$T python tools/s2_rate.py 256:2072 300:2072 400:1184 500:1184 600:592 800:296 1000:296 1500:148 2000:148 2>&1 | grep -v Warning
