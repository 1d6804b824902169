"""bflc_demo_b200 -- a Blackwell-native committee-consensus federated-learning engine.

Capability parity target: iammcy/BFLC-demo (FISCO-BCOS precompiled contract + TF1 client);
see SURVEY.md for the component map and DESIGN.md for the B200-first architecture.
"""
__version__ = "0.1.0"
