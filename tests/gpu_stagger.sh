#!/bin/bash
for sg in 0 9000 36000 144000 1000000; do echo "stagger $sg"; TFHE_B200_STAGGER=$sg python tools/brtime.py 444 1024 2>&1 | tail -2; done
