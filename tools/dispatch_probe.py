import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poppunk_mod_b200 as pk
nat = pk._native; L = nat.lib()
names = {0: "FFMA", 1: "DFMA", 2: "FFMA2", 3: "DFMA + 1 FFMA each", 4: "DFMA + 2 FFMA each"}
for w in (1, 3, 4, 0):
    tf, mhz = ctypes.c_double(), ctypes.c_double()
    nat.check(L.kdejs_peak_flops(0, w, ctypes.byref(tf), ctypes.byref(mhz)))
    print(f"{names[w]:22s}: {tf.value:6.2f} TFLOP/s (DFMA/FFMA-only count)")
