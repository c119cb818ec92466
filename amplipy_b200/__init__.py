"""amplipy_b200 -- B200-native trim -> pileup -> call path of AmpliPy.

Layout: csrc/ (CUDA kernels + C ABI, host BGZF/BAM codec), engine.py (ctypes binding mirroring the reference's
per-read / per-position functions), cli.py (the reference's command line), alnio.py / vcf.py / primers.py /
calling.py (file formats around the path), synth.py (seeded synthetic inputs), dist.py (multi-GPU sharding)."""
__version__ = "0.1.0"
