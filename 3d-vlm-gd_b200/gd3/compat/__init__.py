"""Drop-in modules with the reference's Python signatures.

    from gd3.compat import losses     # utils/losses.py
    from gd3.compat import functions  # utils/functions.py (hot-path subset)
    from gd3.compat import fast_nn    # mast3r/fast_nn.py
"""
