"""Tuning helper: run another script with a variant build of the library (jtransforms_b200.build.build_variant).
usage: python scripts/with_lib.py <libjtb200_xxx.so> <script.py> [args...]"""
import os, sys, runpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jtransforms_b200 import _lib
_lib.use(os.path.abspath(sys.argv[1]))
sys.argv = sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
