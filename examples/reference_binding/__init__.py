"""INTEGRATION.md section B as runnable code: the two modules a maintainer of the reference
would swap in (`reinfocus/graphics/render.py` and `reinfocus/vision.py`), binding the C-ABI of
include/reinfocus_b200.h with plain ctypes. Everything else of the reference - camera / world
host packing, DeviceData, FocusObserver, the env classes - runs unmodified on top of them
(scripts/run_reference_binding.py does exactly that and replays a golden env sequence)."""
