#!/usr/bin/env python
"""Drop-in for SquiggleKit's dRNA_segmenter.py (slow5 branch) running on B200 (see squigglekit_b200/cli_drna_segmenter.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from squigglekit_b200.cli_drna_segmenter import main  # noqa: E402

if __name__ == '__main__':
    main()
