"""CPU oracles for the NNUE hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this package; nothing under nnue-vision_b200/ does (tests/test_layout.py checks).
"""
