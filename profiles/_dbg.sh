for m in 0 15 7 8 4 3; do echo "dbg $m"; VK_DBG=$m VK_CONV_MODE=1 timeout 60 python profiles/conv_head_bench.py 64 2>&1 >/dev/null | grep "yolov5s demo" | cut -c1-40; done
