"""CPU oracle for the LMaze hot path -- test infrastructure only (see lmaze_oracle.h)."""
