"""Generates tests/golden/keypoints_cases.json by running the UNMODIFIED reference `read_json_keypoint`
(PGNR/utils/utils.py:12-60, imported through oracle/ref_import.py) on synthetic OpenPose files.

    python oracle/make_golden_keypoints.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def person(rng, cx, cy, size, conf_body=0.9, n_hand_valid=(21, 21), drop_body=()):
    body = np.zeros((25, 3))
    body[:, 0] = cx + rng.uniform(-size, size, 25)
    body[:, 1] = cy + rng.uniform(-size, size, 25)
    body[:, 2] = conf_body * rng.uniform(0.5, 1.0, 25)
    for i in drop_body:
        body[i] = 0.0
    hands = []
    for nv in n_hand_valid:
        h = np.zeros((21, 3))
        h[:, 0] = cx + rng.uniform(-20, 20, 21)
        h[:, 1] = cy + rng.uniform(-20, 20, 21)
        h[:nv, 2] = rng.uniform(0.2, 1.0, nv)
        hands.append(h)
    return {'pose_keypoints_2d': body.reshape(-1).tolist(), 'hand_left_keypoints_2d': hands[0].reshape(-1).tolist(),
            'hand_right_keypoints_2d': hands[1].reshape(-1).tolist()}


def cases():
    rng = np.random.default_rng(7)
    out = [
        ('single person', {'people': [person(rng, 300, 200, 80)]}),
        ('larger of two people wins', {'people': [person(rng, 100, 100, 30), person(rng, 400, 300, 120)]}),
        ('first person larger', {'people': [person(rng, 400, 300, 150), person(rng, 100, 100, 30)]}),
        ('person with fewer than 4 confident joints is skipped',
         {'people': [person(rng, 300, 200, 200, drop_body=range(0, 13)), person(rng, 200, 200, 40)]}),
        ('no people', {'people': []}),
        ('nobody valid', {'people': [person(rng, 300, 200, 80, conf_body=0.05)]}),
        ('hands with at most 5 valid points give zero rows', {'people': [person(rng, 250, 250, 60, n_hand_valid=(5, 6))]}),
        ('hands without any valid point', {'people': [person(rng, 250, 250, 60, n_hand_valid=(0, 0))]}),
        ('dropped subset joints stay zero', {'people': [person(rng, 250, 250, 60, drop_body=(19, 22, 3))]}),
    ]
    return out


def main():
    ref_import.load()
    from utils.utils import read_json_keypoint
    recs = []
    with tempfile.TemporaryDirectory() as d:
        for name, doc in cases():
            path = os.path.join(d, 'k.json')
            with open(path, 'w') as f:
                json.dump(doc, f)
            recs.append({'name': name, 'json': doc, 'expected': np.asarray(read_json_keypoint(path), dtype=np.float64).tolist()})
    out = os.path.join(ROOT, 'tests', 'golden', 'keypoints_cases.json')
    with open(out, 'w') as f:
        json.dump({'source': 'PGNR/utils/utils.py:12-60 read_json_keypoint, run by oracle/make_golden_keypoints.py', 'cases': recs}, f)
    print('wrote', out, len(recs), 'cases')


if __name__ == '__main__':
    main()
