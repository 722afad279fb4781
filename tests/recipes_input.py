"""Input of the replayed reference recipes (tests/golden/make_recipes_golden.py, tests/test_reference_recipes.py):
a seeded hBN / graphene stack standing in for the reference's hBN_truncated.lammpstrj (absent from the checkout)."""
from pyslice_b200 import synthetic

A, B = 2.5575, 4.2625            # rectangular 4-atom cell of the synthetic stack
NAMES = {5: "B", 6: "C", 7: "N"}


def recipe_trajectory():
    """6 x 4 cells, 2 layers (hBN + graphene), 12 frames, timestep 0.005 ps: box 15.3 x 17.1 x 6.7 A -> 153 x 170 x 14"""
    return synthetic.hbn_graphene_trajectory(cells=(6, 4), n_layers=2, n_frames=12, seed=11, timestep=0.005)
