timeout 120 python tools/stage_time.py c1 12 2>&1 | tail -1
B200SGM_VERT_HALO_FIT=0 timeout 120 python tools/stage_time.py c1 12 2>&1 | tail -1
