import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
warnings.filterwarnings("ignore", category=SyntaxWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
