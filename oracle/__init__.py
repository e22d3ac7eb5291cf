"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/mrts_oracle.h).  Never imported by microrts_b200."""
