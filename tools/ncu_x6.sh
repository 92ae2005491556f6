set -x
export A3D_XATTN_CORE=6
timeout 300 ncu --set full --import-source on --clock-control none -k regex:xattn6 -s 2 -c 1 -o gpurun_out/x6_np6 python tools/xattn_one.py 3 2>&1 | tail -5
