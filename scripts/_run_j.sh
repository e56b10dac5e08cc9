python scripts/phase_timing_gn.py --batch 256 --size 64 2>&1 | tail -9
python scripts/phase_timing_gn.py --batch 128 --size 32 2>&1 | tail -3
