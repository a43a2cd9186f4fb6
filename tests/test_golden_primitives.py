"""The oracle's primitives against digests frozen from the reference's SECOND (encoder-side) copies of the same
primitives (tests/golden/primitives_second_copies.json, made by tools/make_golden_primitives.py where /root/reference
exists): motion compensation at all 16 sizes x 4 phases, 600 coefficient blocks through the inverse transforms, every
directional / DC predictor on luma, U and V, the three plane predictors.  Needs neither the reference nor oracle/_ref."""
import json
import os

import primitive_cases as pc

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'primitives_second_copies.json')


def test_oracle_primitives_match_the_references_second_copies():
    want = json.load(open(PATH))['digests']
    got = pc.digests(pc.run_oracle)
    assert got == want
