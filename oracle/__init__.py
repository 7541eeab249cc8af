"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/kr_oracle.h). Never imported by kuberay_b200/."""
