"""Stub (never called on the hot path)."""
