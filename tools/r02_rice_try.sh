#!/bin/bash
timeout 600 python tools/rice_decode_roofline.py 48 --tiles 1,16 --splits auto --out gpurun_out/_x.json 2>&1 | grep streams
