"""jumanji==0.2.2 jumanji/environments/routing/connector/constants.py (UPSTREAM, restated).
Cell codes are pinned by the reference (seed_extension.py:39-44, grid_utils.py:12), action codes by
test_parallel_random_walk_board.py:361-378."""
EMPTY = 0
PATH = 1
POSITION = 2
TARGET = 3
AGENT_INITIAL_VALUE = 1
NOOP = 0
UP = 1
RIGHT = 2
DOWN = 3
LEFT = 4
