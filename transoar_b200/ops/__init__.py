"""Mirror of ``transoar.models.ops`` (functions/ + modules/)."""
