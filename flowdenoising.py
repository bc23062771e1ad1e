#!/usr/bin/env python
"""Drop-in command line of the reference's src/flowdenoising.py (same flags), running the B200 implementation:

    python flowdenoising.py -i volume.mrc -o denoised_volume.mrc -s 2 2 2 -l 3 -w 5
"""
import sys

from flowdenoising_b200.flowdenoising import main

if __name__ == "__main__":
    sys.exit(main())
