"""Loss-side helpers that sit right next to the op."""
