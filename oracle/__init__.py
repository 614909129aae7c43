"""CPU oracle for the EVP path -- TEST INFRASTRUCTURE ONLY (see oracle/evp_oracle.h)."""
