#!/usr/bin/env python3
"""Audio -> log-mel .npy with the reference's flags (convert_spectrograms.py:91-133), mel extraction on the GPU."""
from mqgan_b200.convert_spectrograms import main

if __name__ == "__main__":
    main()
