"""Import shim so tests can reuse tools/make_golden_orb.py's cv2 wrappers."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "make_golden_orb.py")
_spec = importlib.util.spec_from_file_location("make_golden_orb", _p)
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
orb_canonical = _m.orb_canonical
knn2 = _m.knn2
