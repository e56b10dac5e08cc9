"""Kept for the fixture generator and the live-reference test: the importer lives in baseline/ref_import.py."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from baseline.ref_import import REF_ROOT, import_reference, reference_available  # noqa: E402,F401
