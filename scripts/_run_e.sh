python scripts/phase_timing.py --batch 256 --size 64 2>&1 | tail -22
python scripts/phase_timing.py --batch 128 --size 32 2>&1 | tail -14
python scripts/phase_timing.py --batch 256 --size 32 --cin 128 --cout 128 2>&1 | tail -14
