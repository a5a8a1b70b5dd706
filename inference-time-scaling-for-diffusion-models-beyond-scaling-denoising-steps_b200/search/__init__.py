"""Mirror of the reference's `search` package: search_algorithm.py, verifier.py."""
