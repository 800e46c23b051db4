import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bm = importlib.import_module("automated-deep-photo-style-transfer_b200.benchmark")
print(json.dumps(bm.run(verbose=False)))
