"""TEST INFRASTRUCTURE — CPU oracle (parity unpinned, see oracle_math.h). Not part of the product."""
