#!/usr/bin/env python
"""Drop-in for SquiggleKit's SquigglePull.py (see squigglekit_b200/cli_squigglepull.py): fast5 -> signal TSV, host-side."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from squigglekit_b200.cli_squigglepull import main  # noqa: E402

if __name__ == '__main__':
    main()
