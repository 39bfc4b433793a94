"""Score-network modules with the reference's names, constructor arguments and state-dict keys
(models/ncsnpp.py, models/layerspp.py, models/layers.py, models/up_or_down_sampling.py)."""
