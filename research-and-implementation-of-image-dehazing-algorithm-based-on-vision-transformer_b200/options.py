"""Run-time switches mirrored from the reference's options.py.

``is_relative_position_bias`` is the module-level ablation flag of options.py:5 that
ProbSparse/attn.py:227 imports on every call.
"""
is_relative_position_bias = True
