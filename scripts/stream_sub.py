"""The `streaming` sub-record of bench.py on its own (gpurun: python scripts/stream_sub.py > gpurun_out/x.json)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from aaltoasr_b200 import AkuGpu

from aaltoasr_b200 import synth
eng = AkuGpu(0)
eng.frontend_load_config_text(synth.mfcc39_config(bench.SAMPLE_RATE))
model = bench.make_model(eng)
eng.close()
print(json.dumps(bench.sub_streaming(argparse.Namespace(), 0, model)))
