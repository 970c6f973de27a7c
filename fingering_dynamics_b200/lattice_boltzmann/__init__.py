"""Drop-in twins of the reference's lattice_boltzmann/*.py modules (same module, class and function
names, NumPy arrays in and out), written from scratch on top of the CUDA engine.  The reference scripts
import their siblings by bare name (`from create_block import Createblock`); adding this directory to
sys.path keeps that working, and the modules are also importable as a package."""
