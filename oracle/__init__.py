"""CPU oracle — test infrastructure only (see oracle/schw_oracle.c).  Never imported by the product."""
