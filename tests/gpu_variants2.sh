#!/bin/bash
for v in 3 4 5 6 7; do TFHE_B200_STAGGER=0 TFHE_B200_BR_VARIANT=$v python tools/brtime.py 1024 7104 2>&1 | tail -2; done
