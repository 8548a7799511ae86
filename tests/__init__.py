"""Parity and host-logic tests of gym_comm_b200 (CPU: -m \"not gpu\"; B200: -m gpu)."""
