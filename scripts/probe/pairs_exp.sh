for mp in 74 66 60 50; do echo "max pairs $mp"; MDSCTK_TC_MAX_PAIRS=$mp ITER_TESTS=0 ITER_DBG="3:0" bash scripts/gpu_iter.sh | grep kernel; done
