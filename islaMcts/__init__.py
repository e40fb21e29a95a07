"""Compatibility package: the reference's module paths (islaMcts.agents.mcts_hash.MctsHash, ...) resolve to
the B200 engine's host layer, so code written against LorenzoBonanni/IslaMcts imports unchanged."""
