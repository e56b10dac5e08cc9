python scripts/phase_timing_wgrad.py 2>&1 | tail -2
python scripts/phase_timing_wgrad.py --cin 128 --cout 64 2>&1 | tail -2
python scripts/phase_timing_wgrad.py --size 16 --cin 128 --cout 128 2>&1 | tail -2
