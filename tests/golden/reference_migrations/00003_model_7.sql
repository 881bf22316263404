INSERT INTO models (id, name, model_type, created_at) VALUES
  (7, 'MsMarcoBertBaseDotV5', 'MsMarcoBertBaseDotV5', 0);

INSERT INTO model_versions (model_id, version, status, weights_filename, created_at) VALUES
  (7, 0, 'ready', '', 0);

