"""TEST INFRASTRUCTURE ONLY: CPU oracle for the DEFLATE hot path (see inflate_oracle.c)."""
