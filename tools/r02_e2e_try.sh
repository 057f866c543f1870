#!/bin/bash
python tools/e2e_ab.py SELAB200_DEC_TAPER 0 1 1
python tools/e2e_ab.py SELAB200_DEC_PARTS 12 16 1
SELAB200_DEC_TAPER=1 python tools/e2e_ab.py SELAB200_DEC_PARTS 6 12 1
