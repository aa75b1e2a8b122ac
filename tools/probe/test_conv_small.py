import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import importlib.util
spec = importlib.util.spec_from_file_location('tc', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'test_conv.py'))
src = open(spec.origin).read().replace("for n, N in ((11, 8), (11, 4096), (19, 5), (11, 40960)):", "for n, N in ((11, 8),):")
exec(compile(src, spec.origin, 'exec'))
