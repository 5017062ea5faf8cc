"""CPU oracle of the hot path -- TEST INFRASTRUCTURE ONLY (see the module headers)."""
