import sys, os, torch
sys.path.insert(0, os.getcwd())
torch.autograd.set_detect_anomaly(True)
import __graft_entry__ as g
try:
    g.smoke()
except Exception as e:
    import traceback; traceback.print_exc()
