"""CPU oracle for the hot path (test infrastructure; see the module headers)."""
